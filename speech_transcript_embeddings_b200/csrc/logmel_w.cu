// Recipe W (Whisper log-mel) for sm_100a.
//
// Replaces WhisperFeatureExtractor.__call__ + _torch_extract_fbank_features
// (TF/models/whisper/feature_extraction_whisper.py:135-164, 189-342): zero-pad / truncate to n_samples,
// centre reflect-pad 200, periodic Hann-400, 400-point real DFT every 160 samples (last frame dropped),
// |.|^2, Slaney mel (201 -> 80), log10(max(., 1e-10)), per-clip max(x, max - 8), (x + 4) / 4,
// layout [B, 80, n_samples/160], optional mask = sample mask every 160th sample.
//
// The reference's default path is float32 (torch.stft); this kernel is float32 throughout.
//
//   w_frames    one CTA per (clip, chunk of 128 frames); 8 threads per frame in pass 1
//               z[n] = y[2n] + j y[2n+1], n < 200; 200-point complex FFT as 25 x 8:
//                 pass 1  thread n2 (8 per frame): radix-25 (5 x 5) over z[n2 + 8 n1], times W200^(n2 k1)
//                 pass 2  100 rows (4 frames x 25) per warp: 8-point DFT over n2 -> Z[k1 + 25 k2]
//               real split (pairs k, 200-k), power, sparse mel, log10 -> out (raw) + per-clip max (atomic)
//   w_finish    in-place max(x, max - 8), (x + 4) / 4 and the mask
#include "stx_common.h"
#include <cmath>
#include <mutex>

namespace stx {
namespace {

constexpr int kN = STX_W_NFFT;          // 400
constexpr int kHop = STX_W_HOP;         // 160
constexpr int kMel = STX_W_NMEL;        // 80
constexpr int kBins = kN / 2 + 1;       // 201
constexpr int kThreads = 256;
constexpr int kRound = 32;              // frames in flight per CTA (4 per warp)
constexpr int kChunk = 128;             // frames per CTA
constexpr int kTile = (kRound - 1) * kHop + kN;   // 5360 samples
constexpr int kExRow = 9;               // padded row of 8 complex
constexpr int kExFrame = 232;           // complex per frame in the exchange (25 * 9 = 225, padded: 1856 B = 64 mod 128)
constexpr int kPRow = 203;              // odd stride for the power spectrum rows
constexpr int kMelWeights = 400;        // >= 391 non-zeros

struct WTables {
    float  win[kN];
    float2 tw[25 * 8];          // [k1][n2] = W200^(n2 k1)
    float2 post[104];           // W400^k, k = 0..100
    float  melw[kMelWeights];   // 0.25 * weights
    int    melmeta[kMel];       // first | count << 9 | offset << 18
};

struct Smem {
    float  pcm[kTile];
    float  win[kN];
    float2 tw[200];
    float2 post[104];
    float  melw[kMelWeights];
    int    melmeta[kMel];
    float2 ex[kRound * kExFrame];
    float  P[kRound * kPRow];
    float  wmax[kThreads / 32];
};
static_assert(sizeof(Smem) <= 112 * 1024, "two CTAs per SM must fit");

struct cf { float re, im; };
__device__ __forceinline__ cf operator+(cf a, cf b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cf operator-(cf a, cf b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cf cmul(cf a, float wr, float wi) {
    return {fmaf(a.re, wr, -(a.im * wi)), fmaf(a.re, wi, a.im * wr)};
}
__device__ __forceinline__ cf cfma(float s, cf a, cf b) { return {fmaf(s, a.re, b.re), fmaf(s, a.im, b.im)}; }

// forward 5-point DFT
__device__ __forceinline__ void dft5(cf a0, cf a1, cf a2, cf a3, cf a4, cf& A0, cf& A1, cf& A2, cf& A3, cf& A4) {
    constexpr float c1 = 0.30901699437494742410f;    // cos(2 pi / 5)
    constexpr float c2 = -0.80901699437494742410f;   // cos(4 pi / 5)
    constexpr float s1 = 0.95105651629515357212f;    // sin(2 pi / 5)
    constexpr float s2 = 0.58778525229247312917f;    // sin(4 pi / 5)
    cf t1 = a1 + a4, t2 = a2 + a3, t3 = a1 - a4, t4 = a2 - a3;
    A0 = a0 + t1 + t2;
    cf m1 = cfma(c2, t2, cfma(c1, t1, a0));
    cf m2 = cfma(c1, t2, cfma(c2, t1, a0));
    cf n1 = {fmaf(s2, t4.re, s1 * t3.re), fmaf(s2, t4.im, s1 * t3.im)};
    cf n2 = {fmaf(-s1, t4.re, s2 * t3.re), fmaf(-s1, t4.im, s2 * t3.im)};
    A1 = {m1.re + n1.im, m1.im - n1.re};
    A4 = {m1.re - n1.im, m1.im + n1.re};
    A2 = {m2.re + n2.im, m2.im - n2.re};
    A3 = {m2.re - n2.im, m2.im + n2.re};
}

// W25^e = (cos(2 pi e / 25), -sin(2 pi e / 25)) for the exponents q * k1 that occur (q, k1 in 1..4);
// e is a compile-time constant after unrolling, so the switch folds away
__device__ __forceinline__ cf tw25(cf a, int e) {
    switch (e) {
        case 1: return cmul(a, 0.96858316112863107605f, -0.24868988716485479484f);
        case 2: return cmul(a, 0.87630668004386358394f, -0.48175367410171532345f);
        case 3: return cmul(a, 0.72896862742141155245f, -0.68454710592868861507f);
        case 4: return cmul(a, 0.53582679497899654564f, -0.84432792550201507531f);
        case 6: return cmul(a, 0.06279051952931352654f, -0.99802672842827155897f);
        case 8: return cmul(a, -0.42577929156507271502f, -0.90482705246601946580f);
        case 9: return cmul(a, -0.63742398974868974548f, -0.77051324277578925326f);
        case 12: return cmul(a, -0.99211470131447776488f, -0.12533323356430453588f);
        default: return cmul(a, -0.63742398974868952344f, 0.77051324277578936428f);   // 16
    }
}

// forward 25-point DFT, natural order: n = q + 5 m, k = k1 + 5 k2
__device__ __forceinline__ void dft25(const cf (&v)[25], cf (&o)[25]) {
    cf b[5][5];
#pragma unroll
    for (int q = 0; q < 5; ++q)
        dft5(v[q], v[q + 5], v[q + 10], v[q + 15], v[q + 20], b[q][0], b[q][1], b[q][2], b[q][3], b[q][4]);
#pragma unroll
    for (int q = 1; q < 5; ++q)
#pragma unroll
        for (int k1 = 1; k1 < 5; ++k1) b[q][k1] = tw25(b[q][k1], q * k1);
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1)
        dft5(b[0][k1], b[1][k1], b[2][k1], b[3][k1], b[4][k1], o[k1], o[k1 + 5], o[k1 + 10], o[k1 + 15], o[k1 + 20]);
}

// forward 8-point DFT, natural order
__device__ __forceinline__ void dft8(const cf (&v)[8], cf (&o)[8]) {
    constexpr float h = 0.70710678118654752440f;
    // n = q + 2 m (q = 0, 1; m = 0..3), k = k1 + 4 k2
    cf e0 = v[0] + v[4], e1 = v[0] - v[4], e2 = v[2] + v[6], e3 = v[2] - v[6];
    cf E0 = e0 + e2, E2 = e0 - e2;
    cf E1 = {e1.re + e3.im, e1.im - e3.re}, E3 = {e1.re - e3.im, e1.im + e3.re};
    cf f0 = v[1] + v[5], f1 = v[1] - v[5], f2 = v[3] + v[7], f3 = v[3] - v[7];
    cf F0 = f0 + f2, F2 = f0 - f2;
    cf F1 = {f1.re + f3.im, f1.im - f3.re}, F3 = {f1.re - f3.im, f1.im + f3.re};
    // W8^k1 on the odd half
    cf G1 = {(F1.re + F1.im) * h, (F1.im - F1.re) * h};
    cf G2 = {F2.im, -F2.re};
    cf G3 = {(F3.im - F3.re) * h, -(F3.re + F3.im) * h};
    o[0] = E0 + F0; o[4] = E0 - F0;
    o[1] = E1 + G1; o[5] = E1 - G1;
    o[2] = E2 + G2; o[6] = E2 - G2;
    o[3] = E3 + G3; o[7] = E3 - G3;
}

// float max through integer atomics (works for any sign, destination initialised to -inf)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void w_init_max(float* __restrict__ clip_max, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) clip_max[i] = __int_as_float(0xff800000);
}

__global__ void __launch_bounds__(kThreads, 2)
w_frames(const float* __restrict__ pcm, const long long* __restrict__ offsets, const int* __restrict__ lengths,
         const float* __restrict__ peaks, const WTables* __restrict__ tab, int n_samples, float* __restrict__ out,
         float* __restrict__ clip_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const int b = blockIdx.y;
    const int T = n_samples / kHop;
    const int t_begin = blockIdx.x * kChunk;
    if (t_begin >= T) return;
    const int t_end = min(T, t_begin + kChunk);
    const int len = min(lengths[b], n_samples);
    const float* clip = pcm + offsets[b];
    const float peak = peaks ? peaks[b] : 1.0f;
    float* out_b = out + (size_t)b * kMel * T;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < kN; i += kThreads) sm.win[i] = tab->win[i];
    if (tid < 200) sm.tw[tid] = tab->tw[tid];
    if (tid < 104) sm.post[tid] = tab->post[tid];
    for (int i = tid; i < kMelWeights; i += kThreads) sm.melw[i] = tab->melw[i];
    if (tid < kMel) sm.melmeta[tid] = tab->melmeta[tid];

    float run_max = __int_as_float(0xff800000);

    for (int t0 = t_begin; t0 < t_end; t0 += kRound) {
        __syncthreads();
        // ---- PCM tile with zero padding to n_samples and reflect padding of 200 around it ----
        const int g0 = t0 * kHop - kN / 2;
        for (int i = tid; i < kTile; i += kThreads) {
            int g = g0 + i;
            if (g < 0) g = -g;
            if (g >= n_samples) g = 2 * (n_samples - 1) - g;
            float x = (g >= 0 && g < len) ? __ldg(clip + g) : 0.0f;
            if (peak != 1.0f) x = x / peak;          // float32 division, like numpy's (R/processor.py:92)
            sm.pcm[i] = x;
        }
        __syncthreads();

        // ---- pass 1: 8 threads per frame, 4 frames per warp ----
        {
            const int fw = lane >> 3, n2 = lane & 7;
            const int fr = warp * 4 + fw;
            const float2* x2 = reinterpret_cast<const float2*>(sm.pcm + fr * kHop);
            const float2* w2 = reinterpret_cast<const float2*>(sm.win);
            cf v[25], a[25];
#pragma unroll
            for (int n1 = 0; n1 < 25; ++n1) {
                const float2 x = x2[n2 + 8 * n1], w = w2[n2 + 8 * n1];
                v[n1] = {x.x * w.x, x.y * w.y};
            }
            dft25(v, a);
            float2* ex = sm.ex + fr * kExFrame;
            ex[n2] = make_float2(a[0].re, a[0].im);
#pragma unroll
            for (int k1 = 1; k1 < 25; ++k1) {
                const float2 w = sm.tw[k1 * 8 + n2];
                const cf m = cmul(a[k1], w.x, w.y);
                ex[k1 * kExRow + n2] = make_float2(m.re, m.im);
            }
        }
        __syncwarp();

        // ---- pass 2: the warp's 100 rows (frame, k1), 8-point DFT each ----
        cf zz[4][8];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int row = lane + 32 * it;
            if (row < 100) {
                const int fw = row / 25, k1 = row - fw * 25;
                const float2* ex = sm.ex + (warp * 4 + fw) * kExFrame + k1 * kExRow;
                cf v[8];
#pragma unroll
                for (int n2 = 0; n2 < 8; ++n2) { const float2 p = ex[n2]; v[n2] = {p.x, p.y}; }
                dft8(v, zz[it]);
            }
        }
        __syncwarp();
        // natural order Z[k1 + 25 k2] back into the frame's exchange area
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int row = lane + 32 * it;
            if (row < 100) {
                const int fw = row / 25, k1 = row - fw * 25;
                float2* z = sm.ex + (warp * 4 + fw) * kExFrame;
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) z[k1 + 25 * k2] = make_float2(zz[it][k2].re, zz[it][k2].im);
            }
        }
        __syncwarp();

        // ---- real split + power: pairs (k, 200 - k), k = 0..100, of the warp's 4 frames ----
        for (int item = lane; item < 4 * 101; item += 32) {
            const int fw = item / 101, k = item - fw * 101;
            const int fr = warp * 4 + fw;
            const float2* z = sm.ex + fr * kExFrame;
            const float2 zk = z[k], zp = z[k == 0 ? 0 : 200 - k];
            const float2 w = sm.post[k];
            const float sr = zk.x + zp.x, dr = zk.x - zp.x, si = zk.y + zp.y, di = zk.y - zp.y;
            const float u = fmaf(w.y, dr, w.x * si);
            const float vv = fmaf(w.y, si, -(w.x * dr));
            const float xr = sr + u, xi = di + vv;       // 2 X[k]
            const float yr = sr - u, yi = vv - di;       // 2 X[200 - k]
            float* P = sm.P + fr * kPRow;
            P[k] = fmaf(xr, xr, xi * xi);
            P[200 - k] = fmaf(yr, yr, yi * yi);
        }
        __syncthreads();

        // ---- sparse mel + log10; lane <-> frame so that stores along t coalesce ----
        {
            const int t = t0 + lane;
            const float* P = sm.P + lane * kPRow;
#pragma unroll 2
            for (int m = warp; m < kMel; m += kThreads / 32) {
                const int meta = sm.melmeta[m];
                const int first = meta & 511, count = (meta >> 9) & 511, off = meta >> 18;
                float acc = 0.0f;
                for (int q = 0; q < count; ++q) acc = fmaf(sm.melw[off + q], P[first + q], acc);
                const float lg = log10f(fmaxf(acc, 1e-10f));
                if (t < t_end) {
                    out_b[(size_t)m * T + t] = lg;
                    run_max = fmaxf(run_max, lg);
                }
            }
        }
    }

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) run_max = fmaxf(run_max, __shfl_xor_sync(0xffffffffu, run_max, o));
    if (lane == 0) sm.wmax[warp] = run_max;
    __syncthreads();
    if (tid == 0) {
        float m = sm.wmax[0];
#pragma unroll
        for (int w = 1; w < kThreads / 32; ++w) m = fmaxf(m, sm.wmax[w]);
        atomic_max_float(clip_max + b, m);
    }
}

// max(x, max - 8), (x + 4) / 4 in place; mask[b, j] = (160 j < len)
__global__ void __launch_bounds__(256)
w_finish(const int* __restrict__ lengths, const float* __restrict__ clip_max, int n_samples,
         float* __restrict__ out, int* __restrict__ mask) {
    const int b = blockIdx.y;
    const int T = n_samples / kHop;
    const float lo = clip_max[b] - 8.0f;
    const size_t total = (size_t)kMel * T;
    float* o = out + (size_t)b * total;
    const bool vec = (total % 4 == 0);
    if (vec) {
        float4* o4 = reinterpret_cast<float4*>(o);
        for (size_t q = blockIdx.x * blockDim.x + threadIdx.x; q < total / 4; q += (size_t)gridDim.x * blockDim.x) {
            float4 v = o4[q];
            v.x = (fmaxf(v.x, lo) + 4.0f) / 4.0f;
            v.y = (fmaxf(v.y, lo) + 4.0f) / 4.0f;
            v.z = (fmaxf(v.z, lo) + 4.0f) / 4.0f;
            v.w = (fmaxf(v.w, lo) + 4.0f) / 4.0f;
            o4[q] = v;
        }
    } else {
        for (size_t q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x)
            o[q] = (fmaxf(o[q], lo) + 4.0f) / 4.0f;
    }
    if (mask) {
        const int len = min(lengths[b], n_samples);
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < T; j += gridDim.x * blockDim.x)
            mask[(size_t)b * T + j] = (j * kHop < len) ? 1 : 0;
    }
}

std::mutex g_tab_mutex;
WTables* g_tab[64] = {nullptr};

int get_tables(const WTables** out) {
    int dev = 0;
    STX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return STX_EINVAL; }
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    if (!g_tab[dev]) {
        static WTables h;
        const std::vector<double>& w = w_window();
        for (int i = 0; i < kN; ++i) h.win[i] = (float)w[i];
        for (int k1 = 0; k1 < 25; ++k1)
            for (int n2 = 0; n2 < 8; ++n2) {
                double ang = -2.0 * M_PI * double(n2 * k1) / 200.0;
                h.tw[k1 * 8 + n2] = make_float2((float)std::cos(ang), (float)std::sin(ang));
            }
        for (int k = 0; k < 104; ++k) {
            double ang = -2.0 * M_PI * double(k) / 400.0;
            h.post[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
        MelCsr csr = build_mel_csr(w_mel(), kBins, kMel, 0.25);
        if (csr.weights.size() > size_t(kMelWeights)) { set_error("mel table overflow"); return STX_EINVAL; }
        for (int i = 0; i < kMelWeights; ++i) h.melw[i] = i < int(csr.weights.size()) ? csr.weights[i] : 0.0f;
        for (int m = 0; m < kMel; ++m) h.melmeta[m] = csr.first[m] | (csr.count[m] << 9) | (csr.offset[m] << 18);
        WTables* d = nullptr;
        STX_CUDA(cudaMalloc(&d, sizeof(WTables)));
        STX_CUDA(cudaMemcpy(d, &h, sizeof(WTables), cudaMemcpyHostToDevice));
        STX_CUDA(cudaFuncSetAttribute(w_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        g_tab[dev] = d;
    }
    *out = g_tab[dev];
    return 0;
}

}  // namespace
}  // namespace stx

extern "C" {

int stx_logmel_w_workspace(int B, int n_samples, size_t* bytes) {
    if (B < 0 || n_samples < 0 || !bytes) { stx::set_error("stx_logmel_w_workspace: bad argument"); return STX_EINVAL; }
    *bytes = (size_t(B) * sizeof(float) + 255) & ~size_t(255);
    if (*bytes == 0) *bytes = 256;
    return 0;
}

int stx_logmel_w(const float* d_pcm, const int64_t* d_offsets, const int32_t* d_lengths, int B, int n_samples,
                 const float* d_peak, float* d_out, int32_t* d_mask, void* d_ws, size_t ws_bytes, void* stream) {
    using namespace stx;
    if (B < 0 || n_samples < kN || n_samples % kHop != 0) {
        set_error("stx_logmel_w: need B >= 0 and n_samples >= 400, a multiple of 160 (got %d)", n_samples);
        return STX_EINVAL;
    }
    if (B == 0) return 0;
    if (!d_pcm || !d_offsets || !d_lengths || !d_out || !d_ws) { set_error("stx_logmel_w: null pointer"); return STX_EINVAL; }
    if (B > 65535) { set_error("stx_logmel_w: B = %d > 65535 clips per call", B); return STX_EINVAL; }
    size_t need = 0;
    stx_logmel_w_workspace(B, n_samples, &need);
    if (ws_bytes < need) { set_error("stx_logmel_w: workspace %zu < %zu bytes", ws_bytes, need); return STX_ENOSPACE; }
    if (int rc = check_device()) return rc;
    const WTables* tab = nullptr;
    if (int rc = get_tables(&tab)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* clip_max = static_cast<float*>(d_ws);
    const int T = n_samples / kHop;
    STX_LAUNCH(w_init_max, dim3((B + 255) / 256), dim3(256), 0, st, clip_max, B);
    STX_LAUNCH(w_frames, dim3((T + kChunk - 1) / kChunk, B), dim3(kThreads), sizeof(Smem), st,
               d_pcm, reinterpret_cast<const long long*>(d_offsets), d_lengths, d_peak, tab, n_samples, d_out, clip_max);
    const size_t total = (size_t)kMel * T;
    const int gx = (int)std::max<size_t>(1, std::min<size_t>((total / 4 + 255) / 256, 64));
    STX_LAUNCH(w_finish, dim3(gx, B), dim3(256), 0, st, d_lengths, clip_max, n_samples, d_out, d_mask);
    return 0;
}

}  // extern "C"
