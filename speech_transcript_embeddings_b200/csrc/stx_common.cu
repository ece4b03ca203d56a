// Host-side plumbing of libstx_b200 (see stx_common.h).
#include "stx_common.h"
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <mutex>

namespace stx {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return static_cast<int>(e);
}

// ---- per-launch timing -----------------------------------------------------------------
std::atomic<int> g_profile{0};
namespace {
struct ProfRec { const char* name; cudaEvent_t start, stop; };
std::mutex g_prof_mutex;
std::vector<ProfRec> g_prof;
}  // namespace

void profile_before(const char* name, cudaStream_t st) {
    ProfRec r{name, nullptr, nullptr};
    if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
    cudaEventRecord(r.start, st);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof.push_back(r);
}

void profile_after(cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (!g_prof.empty()) cudaEventRecord(g_prof.back().stop, st);
}

int check_device() {
    static std::atomic<int> cached{1};   // 1 = unknown
    int c = cached.load();
    if (c != 1) return c;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice"); return STX_EDEVICE; }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) { cuda_fail(e, "cudaGetDeviceProperties"); return STX_EDEVICE; }
    if (p.major != 10) {
        set_error("libstx_b200 is built for sm_100a only; device %d is sm_%d%d", dev, p.major, p.minor);
        cached.store(STX_EDEVICE);
        return STX_EDEVICE;
    }
    cached.store(0);
    return 0;
}

// ---- tables ------------------------------------------------------------------------------
// Povey window: (0.5 - 0.5 cos(2 pi i / 399)) ^ 0.85, i = 0..399      (TF/audio_utils.py:593, 601-602)
const std::vector<double>& k_window() {
    static const std::vector<double> w = [] {
        std::vector<double> v(STX_K_FRAME);
        // same odd-integer grid numpy's hanning uses, so the table rounds like the reference's
        for (int i = 0; i < STX_K_FRAME; ++i) {
            double n = double(1 - STX_K_FRAME + 2 * i);
            double hann = 0.5 + 0.5 * std::cos(M_PI * n / double(STX_K_FRAME - 1));
            v[i] = std::pow(hann, 0.85);
        }
        return v;
    }();
    return w;
}

// Periodic Hann(400) = symmetric Hann(401) without its last sample (TF/.../whisper:141, torch.hann_window)
const std::vector<double>& w_window() {
    static const std::vector<double> w = [] {
        std::vector<double> v(STX_W_NFFT);
        for (int i = 0; i < STX_W_NFFT; ++i) {
            double n = double(1 - (STX_W_NFFT + 1) + 2 * i);
            v[i] = 0.5 + 0.5 * std::cos(M_PI * n / double(STX_W_NFFT));
        }
        return v;
    }();
    return w;
}

static std::vector<double> linspace(double a, double b, int n) {
    std::vector<double> v(n);
    double step = (b - a) / double(n - 1);
    for (int i = 0; i < n; ++i) v[i] = a + step * i;   // numpy: start + arange(n) * step
    v[n - 1] = b;
    return v;
}

// triangles max(0, min((f - c_j)/(c_{j+1}-c_j), (c_{j+2} - f)/(c_{j+2}-c_{j+1})))  (TF/audio_utils.py:371-375)
static std::vector<double> triangles(const std::vector<double>& bin_pos, const std::vector<double>& centres, int nmel) {
    const int nb = int(bin_pos.size());
    std::vector<double> fb(size_t(nb) * nmel, 0.0);
    for (int k = 0; k < nb; ++k)
        for (int m = 0; m < nmel; ++m) {
            double down = -(centres[m] - bin_pos[k]) / (centres[m + 1] - centres[m]);
            double up = (centres[m + 2] - bin_pos[k]) / (centres[m + 2] - centres[m + 1]);
            double v = std::min(down, up);
            fb[size_t(k) * nmel + m] = v > 0.0 ? v : 0.0;
        }
    return fb;
}

// Kaldi mel: 1127 ln(1 + f/700); 82 centres linear in mel between mel(20) and mel(8000); FFT bins
// placed in mel space (triangularize_in_mel_space=True); no area norm   (TF/audio_utils.py:282-283, 516-530)
const std::vector<double>& k_mel() {
    static const std::vector<double> fb = [] {
        auto mel = [](double f) { return 1127.0 * std::log(1.0 + f / 700.0); };
        const int nb = STX_K_NFFT / 2 + 1;
        std::vector<double> centres = linspace(mel(20.0), mel(8000.0), STX_K_NMEL + 2);
        std::vector<double> pos(nb);
        const double width = 16000.0 / double((nb - 1) * 2);
        for (int k = 0; k < nb; ++k) pos[k] = mel(width * k);
        return triangles(pos, centres, STX_K_NMEL);
    }();
    return fb;
}

// Slaney mel scale + Slaney area norm, triangles in Hz, 0..8000 Hz (TF/audio_utils.py:285-296, 338-352, 532-535)
const std::vector<double>& w_mel() {
    static const std::vector<double> fb = [] {
        const double logstep = 27.0 / std::log(6.4);
        auto hz2mel = [&](double f) { return f >= 1000.0 ? 15.0 + std::log(f / 1000.0) * logstep : 3.0 * f / 200.0; };
        auto mel2hz = [&](double m) { return m >= 15.0 ? 1000.0 * std::exp((std::log(6.4) / 27.0) * (m - 15.0)) : 200.0 * m / 3.0; };
        const int nb = STX_W_NFFT / 2 + 1;
        std::vector<double> mels = linspace(hz2mel(0.0), hz2mel(8000.0), STX_W_NMEL + 2);
        std::vector<double> hz(mels.size());
        for (size_t i = 0; i < mels.size(); ++i) hz[i] = mel2hz(mels[i]);
        std::vector<double> pos = linspace(0.0, 8000.0, nb);
        std::vector<double> f = triangles(pos, hz, STX_W_NMEL);
        for (int m = 0; m < STX_W_NMEL; ++m) {
            double enorm = 2.0 / (hz[m + 2] - hz[m]);
            for (int k = 0; k < nb; ++k) f[size_t(k) * STX_W_NMEL + m] *= enorm;
        }
        return f;
    }();
    return fb;
}

MelCsr build_mel_csr(const std::vector<double>& fb, int nbins, int nmel, double scale, int pad, int limit) {
    MelCsr c;
    c.first.resize(nmel); c.count.resize(nmel); c.offset.resize(nmel);
    for (int m = 0; m < nmel; ++m) {
        int lo = -1, hi = -1;
        for (int k = 0; k < nbins; ++k)
            if (fb[size_t(k) * nmel + m] != 0.0) { if (lo < 0) lo = k; hi = k; }
        if (lo < 0) { lo = 0; hi = -1; }
        int count = hi - lo + 1;
        if (pad > 1) {
            count = (count + pad - 1) / pad * pad;
            if (count == 0) count = pad;
            if (lo + count > limit) lo = limit - count;      // leading zeros instead of trailing ones
            if (lo < 0) lo = 0;
        }
        c.first[m] = lo; c.count[m] = count; c.offset[m] = int(c.weights.size());
        for (int k = lo; k < lo + count; ++k)
            c.weights.push_back(k < nbins ? float(fb[size_t(k) * nmel + m] * scale) : 0.0f);
    }
    return c;
}

}  // namespace stx

extern "C" {

int stx_abi_version(void) { return STX_ABI_VERSION; }
const char* stx_last_error(void) { return stx::t_error.c_str(); }
uint64_t stx_kernel_launch_count(void) { return stx::g_launches.load(); }

int stx_profile_enable(int on) {
    stx::g_profile.store(on ? 1 : 0);
    return 0;
}

int stx_profile_collect(char* h_names, float* h_ms, int cap) {
    std::lock_guard<std::mutex> lock(stx::g_prof_mutex);
    int n = 0;
    for (auto& r : stx::g_prof) {
        float ms = 0.0f;
        if (r.stop && cudaEventSynchronize(r.stop) == cudaSuccess) cudaEventElapsedTime(&ms, r.start, r.stop);
        if (n < cap && h_names && h_ms) {
            snprintf(h_names + size_t(n) * 32, 32, "%s", r.name);
            h_ms[n] = ms;
            ++n;
        }
        if (r.start) cudaEventDestroy(r.start);
        if (r.stop) cudaEventDestroy(r.stop);
    }
    stx::g_prof.clear();
    return n;
}

int64_t stx_get_table(const char* name, double* h_out, int64_t cap) {
    if (!name || !h_out) { stx::set_error("stx_get_table: null argument"); return STX_EINVAL; }
    const std::vector<double>* t = nullptr;
    if (!strcmp(name, "k_window")) t = &stx::k_window();
    else if (!strcmp(name, "k_mel")) t = &stx::k_mel();
    else if (!strcmp(name, "w_window")) t = &stx::w_window();
    else if (!strcmp(name, "w_mel")) t = &stx::w_mel();
    else { stx::set_error("stx_get_table: unknown table '%s'", name); return STX_EINVAL; }
    if (cap < int64_t(t->size())) { stx::set_error("stx_get_table: cap %lld < %zu", (long long)cap, t->size()); return STX_EINVAL; }
    memcpy(h_out, t->data(), t->size() * sizeof(double));
    return int64_t(t->size());
}

}  // extern "C"
