// Device-side polyphase resampling to the model rate: the step BEFORE the log-mel path.
//
// Replaces librosa.resample(audio_array, orig_sr=orig_sr, target_sr=16000) at R/processor.py:82-86 with the arithmetic of
// librosa's res_type="polyphase", i.e. scipy.signal.resample_poly(y, up, down) (scipy/signal/_signaltools.py, defaults
// window=("kaiser", 5.0), padtype="constant") + fix_length to ceil(n * up / down): a Kaiser-windowed sinc low-pass of
// 20 * max(up, down) + 1 taps at the up-sampled rate, designed in float64 on the host, rounded to float32 and scaled by
// `up` in float32 exactly as scipy does, applied as a polyphase filter bank (bank[phase][j] = h_pad[phase + j * up]):
//
//     t = m + n_pre_remove;  p = t * down;  i0 = p / up;  phase = p % up;   y[m] = sum_j bank[phase][j] * x[i0 - j]
//
// (the reference's DEFAULT res_type is "soxr_hq", computed inside the third-party soxr library that is not available
// offline: see oracle/resample.py for what is and is not pinned).
//
// Two kernels, both HBM-streaming with the input window of a tile staged in shared memory:
//   rs_decim<DOWN>  up == 1 (48 k -> 16 k, 32 k -> 16 k, 96 k, 64 k): every output uses the same 20 * DOWN + 1 taps.  A thread
//                   computes 4 consecutive outputs from one register window of the input, with the taps broadcast from
//                   shared memory as float4: 1 + DOWN shared loads per 4 x 4 FMAs instead of 2 per FMA
//   rs_poly         any up / down: a thread owns four outputs of equal phase (m, m + up, ...), filter bank in shared memory
//                   with float4 tap loads
// The per-clip max |y| (the peak-normalise of R/processor.py:91-92 needs it next) is reduced in the same pass.
#include "stx_common.h"
#include <cmath>
#include <map>
#include <mutex>
#include <numeric>
#include <utility>

namespace stx {
namespace {

constexpr int kThreads = 256;
constexpr int kXsMax = 6400;             // floats of staged input per CTA (25 KB: several CTAs per SM overlap staging and filtering)
constexpr int kBankSmemMax = 18432;      // floats of filter bank kept in shared memory (72 KB); larger banks are read through L1

struct Plan {
    int up, down, J, Jp, n_pre_remove, L;
    std::vector<float> h_pad;            // what scipy hands to upfirdn (float32, scaled by up, front-padded)
    float* d_bank = nullptr;             // [up][Jp], device
};

// I0(x) by its power series in long double (every term positive: no cancellation)
long double bessel_i0(long double x) {
    long double sum = 1.0L, term = 1.0L;
    const long double q = x * x / 4.0L;
    for (int k = 1; k < 500; ++k) {
        term *= q / ((long double)k * (long double)k);
        sum += term;
        if (term < sum * 1e-22L) break;
    }
    return sum;
}

void design(int up, int down, Plan& pl) {
    const int max_rate = std::max(up, down);
    const double f_c = 1.0 / max_rate;
    const int half_len = 10 * max_rate;
    const int numtaps = 2 * half_len + 1;
    const double alpha = 0.5 * (numtaps - 1);
    std::vector<double> h(numtaps);
    const long double i0b = bessel_i0(5.0L);
    double sum = 0.0;
    for (int i = 0; i < numtaps; ++i) {
        const double m = i - alpha;
        const double xs = f_c * m;
        const double sinc = xs == 0.0 ? 1.0 : std::sin(M_PI * xs) / (M_PI * xs);
        const double r = (i - alpha) / alpha;
        const double win = (double)(bessel_i0(5.0L * std::sqrt(std::max(0.0L, 1.0L - (long double)r * r))) / i0b);
        h[i] = f_c * sinc * win;
        sum += h[i];
    }
    // numpy's pairwise sum and this sequential one differ by ~1e-16 relative: invisible after the float32 rounding below
    const int n_pre_pad = down - half_len % down;
    pl.up = up; pl.down = down;
    pl.n_pre_remove = (half_len + n_pre_pad) / down;
    pl.L = n_pre_pad + numtaps;
    pl.h_pad.assign(pl.L, 0.0f);
    for (int i = 0; i < numtaps; ++i) pl.h_pad[n_pre_pad + i] = (float)(h[i] / sum) * (float)up;
    pl.J = (pl.L + up - 1) / up;
    pl.Jp = 4 * (((pl.J + 3) / 4) | 1);  // rows of 4 (2 i + 1) floats: float4 loads, and rows of different phases start in
                                         // different 16-byte bank groups
}

std::mutex g_plan_mutex;
std::map<std::tuple<int, int, int>, Plan> g_plans;     // (device, up, down)

int get_plan(int up, int down, const Plan** out) {
    int dev = 0;
    STX_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    auto key = std::make_tuple(dev, up, down);
    auto it = g_plans.find(key);
    if (it == g_plans.end()) {
        Plan pl;
        design(up, down, pl);
        std::vector<float> bank((size_t)up * pl.Jp, 0.0f);
        for (int ph = 0; ph < up; ++ph)
            for (int j = 0; j < pl.J; ++j) {
                const long long k = ph + (long long)j * up;
                if (k < pl.L) bank[(size_t)ph * pl.Jp + j] = pl.h_pad[k];
            }
        STX_CUDA(cudaMalloc(&pl.d_bank, bank.size() * sizeof(float)));
        STX_CUDA(cudaMemcpy(pl.d_bank, bank.data(), bank.size() * sizeof(float), cudaMemcpyHostToDevice));
        it = g_plans.emplace(key, std::move(pl)).first;
    }
    *out = &it->second;
    return 0;
}

__device__ __forceinline__ void clip_peak(float m, float* peaks, int b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    // non-negative floats order like their bit patterns; peaks[] starts at 1.0f, the divisor for clips that never exceed it
    if ((threadIdx.x & 31) == 0 && m > 1.0f) atomicMax(reinterpret_cast<int*>(peaks + b), __float_as_int(m));
}

__global__ void rs_peak_init(float* __restrict__ peaks, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) peaks[i] = 1.0f;
}

// stage x[i_lo .. i_lo + span) of the clip into shared memory, zeros outside [0, n).  Eight independent loads per thread are
// in flight at a time (the staging of a tile is latency-bound: ~24 loads per thread from HBM)
__device__ __forceinline__ void stage_input(float* xs, const float* __restrict__ clip, long long i_lo, int span, int n) {
    for (int k0 = threadIdx.x; k0 < span; k0 += 8 * kThreads) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * kThreads;
            const long long i = i_lo + k;
            v[u] = (k < span && i >= 0 && i < n) ? __ldg(clip + i) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * kThreads;
            if (k < span) xs[k] = v[u];
        }
    }
}

// ---- general up / down --------------------------------------------------------------------------
// Outputs m and m + up have the SAME phase and inputs exactly `down` samples apart, so a thread owns kR = 4 outputs
// m, m + up, m + 2 up, m + 3 up: every float4 of taps (one LDS.128; rows of the bank are Jp = 4 (2 i + 1) floats long, so
// lanes of different phases hit different bank groups) feeds 16 FMAs, against one tap load per FMA when every output
// fetches its own taps.  A CTA tile is kR * up * Q consecutive outputs (Q = groups of `up` base outputs).
constexpr int kR = 4;
template <bool kBankSmem>
__global__ void __launch_bounds__(kThreads)
rs_poly(const float* __restrict__ in, const long long* __restrict__ in_off, const int* __restrict__ in_len,
        const long long* __restrict__ out_off, const int* __restrict__ out_len, const float* __restrict__ bank,
        int up, int down, int J4, int Jp, int n_pre_remove, int Q, float* __restrict__ out, float* __restrict__ peaks) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;
    float* bank_s = smem + kXsMax;
    const int b = blockIdx.y;
    const int n_out = out_len[b];
    const int out_tile = kR * up * Q;
    const int m0 = blockIdx.x * out_tile;
    if (m0 >= n_out) return;
    const int m1 = min(n_out, m0 + out_tile);
    const int n = in_len[b];
    const float* clip = in + in_off[b];
    const int J = 4 * J4;                                              // taps per phase, padded with zeros to a multiple of 4
    const long long i_lo = ((long long)(m0 + n_pre_remove) * down) / up - (J - 1);
    const long long i_hi = ((long long)(m0 + out_tile - 1 + n_pre_remove) * down) / up;      // whole tile: companions may lie past m1
    stage_input(xs, clip, i_lo, (int)(i_hi - i_lo + 1), n);
    if (kBankSmem)
        for (int k = threadIdx.x; k < up * Jp; k += kThreads) bank_s[k] = __ldg(bank + k);
    __syncthreads();
    const float* bk = kBankSmem ? bank_s : bank;
    float* o = out + out_off[b];
    float peak = 0.0f;
    for (int w = threadIdx.x; w < up * Q; w += kThreads) {
        const int q = w / up, base = w - q * up;
        const int m = m0 + q * kR * up + base;                         // this thread's outputs: m + k * up, k = 0 .. kR - 1
        if (m >= m1) continue;
        const long long p = (long long)(m + n_pre_remove) * down;
        const long long i0 = p / up;
        const int ph = (int)(p - i0 * up);
        const float4* hb = reinterpret_cast<const float4*>(bk + (size_t)ph * Jp);
        const float* xr = xs + (int)(i0 - i_lo);
        float acc[kR] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 2
        for (int j4 = 0; j4 < J4; ++j4) {
            const float4 h = hb[j4];
#pragma unroll
            for (int k = 0; k < kR; ++k) {
                const float* x = xr + k * down - 4 * j4;
                acc[k] = fmaf(h.x, x[0], acc[k]);
                acc[k] = fmaf(h.y, x[-1], acc[k]);
                acc[k] = fmaf(h.z, x[-2], acc[k]);
                acc[k] = fmaf(h.w, x[-3], acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < kR; ++k)
            if (m + k * up < m1) { o[m + k * up] = acc[k]; peak = fmaxf(peak, fabsf(acc[k])); }
    }
    if (peaks) clip_peak(peak, peaks, b);
}

// ---- up == 1: integer decimation ------------------------------------------------------------------
// y[m] = sum_k h[k] x[(m + npr) D - k], k = 0 .. L - 1 (L = 21 D + 1 - (10 D) % D... whatever the padded length is).
// A thread owns outputs m .. m + 3.  With the taps taken four at a time (k = 4 q .. 4 q + 3) the inputs it needs for that
// group are the 3 D + 4 consecutive samples x[(m + npr) D - 4 q - 3 .. (m + 3 + npr) D - 4 q]: a sliding register window.
template <int D>
__global__ void __launch_bounds__(kThreads, D <= 3 ? 3 : D == 4 ? 2 : 1)
rs_decim(const float* __restrict__ in, const long long* __restrict__ in_off, const int* __restrict__ in_len,
         const long long* __restrict__ out_off, const int* __restrict__ out_len, const float* __restrict__ taps,
         int n_pre_remove, int out_tile, float* __restrict__ out, float* __restrict__ peaks) {
    constexpr int L = 21 * D + 1;                                      // n_pre_pad (= D: 10 D is a multiple of D) + 20 D + 1
    constexpr int L4 = (L + 3) / 4, Lp = 4 * L4;
    constexpr int W = 3 * D + 4;                                       // register window
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;
    float4* h4 = reinterpret_cast<float4*>(smem + kXsMax);          // taps padded with zeros to Lp
    const int b = blockIdx.y;
    const int n_out = out_len[b];
    const int m0 = blockIdx.x * out_tile;
    if (m0 >= n_out) return;
    const int m1 = min(n_out, m0 + out_tile);
    const int n = in_len[b];
    const float* clip = in + in_off[b];
    const long long i_lo = (long long)(m0 + n_pre_remove) * D - (Lp - 1);
    const long long i_hi = (long long)(((m1 - m0 + 3) & ~3) + m0 - 1 + n_pre_remove) * D;   // whole groups of 4 outputs
    stage_input(xs, clip, i_lo, (int)(i_hi - i_lo + 1), n);
    for (int k = threadIdx.x; k < Lp; k += kThreads) reinterpret_cast<float*>(h4)[k] = k < L ? __ldg(taps + k) : 0.0f;
    __syncthreads();
    float* o = out + out_off[b];
    float peak = 0.0f;
    for (int m = m0 + 4 * threadIdx.x; m < m1; m += 4 * kThreads) {
        // xr[0] = x[(m + npr) D]: output r (0..3) and tap k read xr[r D - k].  (m - m0) D and Lp - 4 are multiples of 4, so
        // xr - 3 is 16-byte aligned and the window moves in float4 steps
        const float* xr = xs + ((m - m0) * D + (Lp - 1));
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        float w[W];                                                    // w[i] = xr[i - 4 q - 3], i = 0 .. 3 D + 3
#pragma unroll
        for (int i = 0; i + 4 <= W; i += 4) {
            const float4 v = *reinterpret_cast<const float4*>(xr + i - 3);
            w[i] = v.x; w[i + 1] = v.y; w[i + 2] = v.z; w[i + 3] = v.w;
        }
#pragma unroll
        for (int i = W & ~3; i < W; ++i) w[i] = xr[i - 3];
#pragma unroll
        for (int q = 0; q < L4; ++q) {
            const float4 h = h4[q];                                    // broadcast
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[r] = fmaf(h.x, w[r * D + 3], acc[r]);
                acc[r] = fmaf(h.y, w[r * D + 2], acc[r]);
                acc[r] = fmaf(h.z, w[r * D + 1], acc[r]);
                acc[r] = fmaf(h.w, w[r * D + 0], acc[r]);
            }
            if (q + 1 < L4) {                                          // slide the window 4 samples towards the past
#pragma unroll
                for (int i = W - 1; i >= 4; --i) w[i] = w[i - 4];
                const float4 v = *reinterpret_cast<const float4*>(xr - 4 * (q + 1) - 3);
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (m + r < m1) { o[m + r] = acc[r]; peak = fmaxf(peak, fabsf(acc[r])); }
    }
    if (peaks) clip_peak(peak, peaks, b);
}

template <int D>
int launch_decim(const Plan& pl, const float* d_taps, dim3 grid, int out_tile, size_t smem, cudaStream_t st,
                 const float* in, const long long* in_off, const int* in_len, const long long* out_off, const int* out_len,
                 float* out, float* peaks) {
    static bool attr_done[64] = {false};
    int dev = 0;
    STX_CUDA(cudaGetDevice(&dev));
    if (!attr_done[dev & 63]) {
        STX_CUDA(cudaFuncSetAttribute(rs_decim<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((kXsMax + 1024) * sizeof(float))));
        attr_done[dev & 63] = true;
    }
    if (pl.L != 21 * D + 1) { set_error("stx_resample_poly: internal: filter length %d != %d", pl.L, 21 * D + 1); return STX_EINVAL; }
    STX_LAUNCH(rs_decim<D>, grid, dim3(kThreads), smem, st, in, in_off, in_len, out_off, out_len, d_taps,
               pl.n_pre_remove, out_tile, out, peaks);
    return 0;
}

}  // namespace
}  // namespace stx

extern "C" {

int stx_resample_plan(int orig_sr, int target_sr, int* up, int* down, int* taps_per_phase, int* n_pre_remove, int* filter_len) {
    using namespace stx;
    if (orig_sr <= 0 || target_sr <= 0) { set_error("stx_resample_plan: sampling rates must be positive"); return STX_EINVAL; }
    const int g = std::gcd(orig_sr, target_sr);
    const int u = target_sr / g, d = orig_sr / g;
    if ((long long)20 * std::max(u, d) + d + 2 > (1 << 22)) { set_error("stx_resample_plan: %d -> %d Hz needs a filter of more than 4 M taps", orig_sr, target_sr); return STX_EINVAL; }
    Plan pl;
    design(u, d, pl);
    if (up) *up = u;
    if (down) *down = d;
    if (taps_per_phase) *taps_per_phase = pl.J;
    if (n_pre_remove) *n_pre_remove = pl.n_pre_remove;
    if (filter_len) *filter_len = pl.L;
    return 0;
}

long long stx_resample_filter(int orig_sr, int target_sr, float* host_out, long long capacity) {
    using namespace stx;
    if (orig_sr <= 0 || target_sr <= 0 || !host_out) { set_error("stx_resample_filter: bad argument"); return STX_EINVAL; }
    const int g = std::gcd(orig_sr, target_sr);
    Plan pl;
    design(target_sr / g, orig_sr / g, pl);
    if (capacity < pl.L) { set_error("stx_resample_filter: capacity %lld < %d", capacity, pl.L); return STX_ENOSPACE; }
    for (int i = 0; i < pl.L; ++i) host_out[i] = pl.h_pad[i];
    return pl.L;
}

int stx_resample_poly(const float* d_in, const int64_t* d_in_offsets, const int32_t* d_in_lengths, int B, int orig_sr,
                      int target_sr, float* d_out, const int64_t* d_out_offsets, const int32_t* d_out_lengths,
                      int max_out_length, float* d_peaks, void* stream) {
    using namespace stx;
    if (B < 0 || orig_sr <= 0 || target_sr <= 0 || max_out_length < 0) { set_error("stx_resample_poly: bad argument"); return STX_EINVAL; }
    if (B == 0) return 0;
    if (!d_in || !d_in_offsets || !d_in_lengths || !d_out || !d_out_offsets || !d_out_lengths) { set_error("stx_resample_poly: null pointer"); return STX_EINVAL; }
    if (orig_sr == target_sr) { set_error("stx_resample_poly: orig_sr == target_sr (nothing to resample; pass the clips through)"); return STX_EINVAL; }
    if (B > 65535) { set_error("stx_resample_poly: B = %d > 65535 clips per call", B); return STX_EINVAL; }
    if (int rc = check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int g = std::gcd(orig_sr, target_sr);
    const int up = target_sr / g, down = orig_sr / g;
    if ((long long)20 * std::max(up, down) + down + 2 > (1 << 22)) { set_error("stx_resample_poly: %d -> %d Hz needs a filter of more than 4 M taps", orig_sr, target_sr); return STX_EINVAL; }
    const Plan* pl = nullptr;
    if (int rc = get_plan(up, down, &pl)) return rc;
    if (d_peaks) STX_LAUNCH(rs_peak_init, dim3((B + 255) / 256), dim3(256), 0, st, d_peaks, B);
    if (max_out_length == 0) return 0;
    const long long* in_off = reinterpret_cast<const long long*>(d_in_offsets);
    const long long* out_off = reinterpret_cast<const long long*>(d_out_offsets);

    if (up == 1 && (down == 2 || down == 3 || down == 4 || down == 6)) {
        const int L4 = (pl->L + 3) / 4;                           // bank row 0 IS h_pad (Jp >= L)
        int out_tile = ((kXsMax - 4 * L4 - 8) / down) & ~3;
        out_tile = std::min(out_tile, 4 * kThreads);
        const dim3 grid((max_out_length + out_tile - 1) / out_tile, B);
        const size_t smem = (size_t)(kXsMax + 4 * L4) * sizeof(float);
        switch (down) {
            case 2: return launch_decim<2>(*pl, pl->d_bank, grid, out_tile, smem, st, d_in, in_off, d_in_lengths, out_off, d_out_lengths, d_out, d_peaks);
            case 3: return launch_decim<3>(*pl, pl->d_bank, grid, out_tile, smem, st, d_in, in_off, d_in_lengths, out_off, d_out_lengths, d_out, d_peaks);
            case 4: return launch_decim<4>(*pl, pl->d_bank, grid, out_tile, smem, st, d_in, in_off, d_in_lengths, out_off, d_out_lengths, d_out, d_peaks);
            default: return launch_decim<6>(*pl, pl->d_bank, grid, out_tile, smem, st, d_in, in_off, d_in_lengths, out_off, d_out_lengths, d_out, d_peaks);
        }
    }

    // general case: a tile is kR * up * Q outputs whose input window must fit the staging buffer
    const int J4 = (pl->J + 3) / 4;
    long long Q = ((long long)(kXsMax - 4 * J4 - 4) * up / down) / ((long long)kR * up);
    Q = std::min<long long>(Q, std::max<long long>(1, (3 * kThreads) / up));       // about three work items per thread
    if (Q < 1) { set_error("stx_resample_poly: %d -> %d Hz: one tile of %d outputs does not fit the staging buffer", orig_sr, target_sr, kR * up); return STX_EINVAL; }
    const long long out_tile = (long long)kR * up * Q;
    const dim3 grid((unsigned)((max_out_length + out_tile - 1) / out_tile), B);
    const bool bank_smem = (long long)up * pl->Jp <= kBankSmemMax;
    static bool attr_done[64] = {false};
    int dev = 0;
    STX_CUDA(cudaGetDevice(&dev));
    if (!attr_done[dev & 63]) {
        STX_CUDA(cudaFuncSetAttribute(rs_poly<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((kXsMax + kBankSmemMax) * sizeof(float))));
        STX_CUDA(cudaFuncSetAttribute(rs_poly<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kXsMax * sizeof(float))));
        attr_done[dev & 63] = true;
    }
    if (bank_smem) {
        STX_LAUNCH(rs_poly<true>, grid, dim3(kThreads), (size_t)(kXsMax + up * pl->Jp) * sizeof(float), st, d_in, in_off,
                   d_in_lengths, out_off, d_out_lengths, pl->d_bank, up, down, J4, pl->Jp, pl->n_pre_remove, (int)Q,
                   d_out, d_peaks);
    } else {
        STX_LAUNCH(rs_poly<false>, grid, dim3(kThreads), (size_t)kXsMax * sizeof(float), st, d_in, in_off, d_in_lengths,
                   out_off, d_out_lengths, pl->d_bank, up, down, J4, pl->Jp, pl->n_pre_remove, (int)Q, d_out, d_peaks);
    }
    return 0;
}

}  // extern "C"
