// Recipe K (Kaldi-style fbank + CMVN + stride-2 stacking) for sm_100a.
//
// Replaces SeamlessM4TFeatureExtractor.__call__ (TF/models/seamless_m4t/feature_extraction_seamless_m4t.py
// :141-302) and the per-frame loop of transformers.audio_utils.spectrogram (TF/audio_utils.py:788-830),
// as reached from R/processor.py:101-105 and R/training/trainer_unfreeze.py:856-860.
//
// Numerics: the reference runs the frame chain and the FFT in float64 and the parity
// bar is a max-abs error on log energies; mel bins that hold 1-2 FFT bins at pre-emphasis-attenuated
// frequencies have chi-square(2) energies, so over 10^5 frames some are 10^-6 of the mean and any
// float32 noise floor (6e-8 of the frame RMS) shows up as >1e-4 in the log.  The frame chain and the
// FFT therefore run on the FP64 pipe; everything after the power spectrum is float32 like the
// reference's own rounding points (complex64 spectrum, float32 log-mel).
//
// Three kernels:
//   k_frames    one CTA per (clip, chunk of 128 frames): PCM -> raw log-mel (written in place into the
//               output tensor, whose [T_pad/2,160] rows are exactly [T_pad,80] rows) + per-chunk
//               per-bin (sum, sum of squares) partials in float64
//   k_finalize  one CTA per clip: ordered reduction of the partials -> mean, 1/sqrt(var_ddof1 + 1e-7)
//   k_normalize in-place CMVN, padding rows, attention mask
//
// Frame pipeline (16 threads per frame, 16 frames in flight per CTA):
//   y[i] = w[i] * (d[i] - c),  d[i] = x[i] - 0.97 x[i-1],  c = 0.03 * mean(frame)      (w[0] = w[399] = 0)
//   z[n] = y[2n] + j y[2n+1]  (n < 200, zero to 256)  -> 256-point complex FFT as 16 x 16:
//     pass 1  thread r:  16-point DFT over z[r + 16 j] (j >= 13 are zero), times W256^(r k1) -> smem
//     pass 2  thread k1: 16-point DFT over r -> Z[k1 + 16 k2]
//   real split  2 X[k] = (Z[k] + conj Z[256-k]) - j W512^k (Z[k] - conj Z[256-k])   (partner by shuffle)
//   power (after rounding X to float32, like the reference's complex64) -> sparse mel -> ln
#include "stx_common.h"
#include <cmath>
#include <cstddef>
#include <mutex>

namespace stx {
namespace {

constexpr int kFrame = STX_K_FRAME;
constexpr int kHop = STX_K_HOP;
constexpr int kMel = STX_K_NMEL;
constexpr int kSlots = 16;                       // frames in flight per CTA
constexpr int kThreads = kSlots * 16;            // 256
constexpr int kChunk = 128;                      // frames per CTA
constexpr int kWinPad = 416;                     // window zero-padded so that r + 16 j <= 207 stays in range
constexpr int kSubSamples = (kSlots - 1) * kHop + kWinPad;   // 2816 = 11 * 256
constexpr int kMelWeights = 768;                 // 501 non-zeros, every filter padded to a multiple of 4
constexpr float kMelFloor = 1.192092955078125e-07f;
constexpr int kStageLead = 4;                    // staging keeps 4 samples before the sub-tile (x[-1] and 16-byte alignment)
constexpr int kStage = kSubSamples + kStageLead; // 2820 floats
constexpr int kSlotFloats = 1024;                // one slot's exchange area (16 x 16 complex doubles) in floats
// power spectrum of frame f starts p_skew(f) floats into its slot: the 16 rows fall on distinct banks for
// the lane <-> frame reads of the mel stage, and the two slots of a warp are 16 banks apart for the stores
__host__ __device__ constexpr int p_skew(int f) { return (f >> 1) + 16 * (f & 1); }
constexpr int kOutRow = kSlotFloats + 17;        // staged log-mel row of frame f at 512 + 1041 f: banks (17 f + m) % 32

struct KTables {
    double  win[kWinPad];        // 2^15 * Povey, zero beyond 400
    double2 tw[16 * 16];         // [k1][r]  = W256^(r k1)
    double2 post[16 * 16];       // [k2][k1] = W512^(k1 + 16 k2) = (cos, -sin)
    float   melw[kMelWeights];   // 0.25 * weights, packed per mel bin
    int     melmeta[kMel];       // first | (count / 4) << 9 | (offset / 4) << 16
};

struct Smem {
    double  dtile[kSubSamples];  // d[i] = x[i] - 0.97 x[i-1] of the current sub-tile (float64)
    double  win[kWinPad];
    double2 tw[256];
    double2 post[256];
    double2 ex[kSlots * 256];    // 16 x 16 exchange per slot (XOR-swizzled); later aliased by the power
                                 // spectra, the staged log-mel rows and the statistics reduction
    float   stage[kStage + 12];  // raw PCM of the NEXT sub-tile, landed by cp.async.bulk (TMA) while this one computes
    float   melw[kMelWeights];
    int     melmeta[kMel];
    double  xb0[kSlots];         // x[160 f]       of each frame of the sub-tile
    double  xb1[kSlots];         // x[160 f + 399]
    unsigned long long mbar;     // completion barrier of the bulk copy
};
static_assert(sizeof(Smem) <= 114 * 1024 - 512, "two CTAs per SM must fit in 228 KB");
static_assert(offsetof(Smem, stage) % 16 == 0 && offsetof(Smem, ex) % 16 == 0, "bulk-copy / vector alignment");

struct cd { double re, im; };
__device__ __forceinline__ cd operator+(cd a, cd b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cd operator-(cd a, cd b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cd cmul(cd a, double wr, double wi) {
    return {fma(a.re, wr, -(a.im * wi)), fma(a.re, wi, a.im * wr)};
}

// forward 4-point DFT (W4 = -j)
__device__ __forceinline__ void dft4(cd a0, cd a1, cd a2, cd a3, cd& A0, cd& A1, cd& A2, cd& A3) {
    cd t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = a1 - a3;
    A0 = t0 + t2;
    A2 = t0 - t2;
    A1 = {t1.re + t3.im, t1.im - t3.re};
    A3 = {t1.re - t3.im, t1.im + t3.re};
}
// same with a3 == 0
__device__ __forceinline__ void dft4_3(cd a0, cd a1, cd a2, cd& A0, cd& A1, cd& A2, cd& A3) {
    cd t0 = a0 + a2, t1 = a0 - a2;
    A0 = t0 + a1;
    A2 = t0 - a1;
    A1 = {t1.re + a1.im, t1.im - a1.re};
    A3 = {t1.re - a1.im, t1.im + a1.re};
}

// forward 16-point DFT, natural order in and out: n = q + 4 m, k = k1 + 4 k2.
// kPruned: inputs 13, 14, 15 are zero (and not read).
template <bool kPruned>
__device__ __forceinline__ void dft16(const cd (&v)[16], cd (&o)[16]) {
    constexpr double c8 = 0.92387953251128675613;   // cos(pi/8)
    constexpr double s8 = 0.38268343236508977173;   // sin(pi/8)
    constexpr double h = 0.70710678118654752440;
    cd b[4][4];
    dft4(v[0], v[4], v[8], v[12], b[0][0], b[0][1], b[0][2], b[0][3]);
    if (kPruned) {
        dft4_3(v[1], v[5], v[9], b[1][0], b[1][1], b[1][2], b[1][3]);
        dft4_3(v[2], v[6], v[10], b[2][0], b[2][1], b[2][2], b[2][3]);
        dft4_3(v[3], v[7], v[11], b[3][0], b[3][1], b[3][2], b[3][3]);
    } else {
        dft4(v[1], v[5], v[9], v[13], b[1][0], b[1][1], b[1][2], b[1][3]);
        dft4(v[2], v[6], v[10], v[14], b[2][0], b[2][1], b[2][2], b[2][3]);
        dft4(v[3], v[7], v[11], v[15], b[3][0], b[3][1], b[3][2], b[3][3]);
    }
    // W16^(q k1)
    b[1][1] = cmul(b[1][1], c8, -s8);                                            // W16^1
    b[1][2] = {(b[1][2].re + b[1][2].im) * h, (b[1][2].im - b[1][2].re) * h};    // W16^2
    b[1][3] = cmul(b[1][3], s8, -c8);                                            // W16^3
    b[2][1] = {(b[2][1].re + b[2][1].im) * h, (b[2][1].im - b[2][1].re) * h};    // W16^2
    b[2][2] = {b[2][2].im, -b[2][2].re};                                         // W16^4 = -j
    b[2][3] = {(b[2][3].im - b[2][3].re) * h, -(b[2][3].re + b[2][3].im) * h};   // W16^6
    b[3][1] = cmul(b[3][1], s8, -c8);                                            // W16^3
    b[3][2] = {(b[3][2].im - b[3][2].re) * h, -(b[3][2].re + b[3][2].im) * h};   // W16^6
    b[3][3] = cmul(b[3][3], -c8, s8);                                            // W16^9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
        dft4(b[0][k1], b[1][k1], b[2][k1], b[3][k1], o[k1], o[k1 + 4], o[k1 + 8], o[k1 + 12]);
}

// ln(x) for normal positive x (the mel floor guarantees it): exponent + MUFU.LG2 of the mantissa.
// |error| <= ~0.6 ulp of the result for results of magnitude 10..30 (the mantissa's log2 is in [0, 1), where
// lg2.approx is accurate to 2^-22 absolute), i.e. as good as logf at a third of the instructions.
__device__ __forceinline__ float ln_pos(float x) {
    const int bits = __float_as_int(x);
    const float e = (float)((bits >> 23) - 127);
    const float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
    float l2;
    asm("lg2.approx.f32 %0, %1;" : "=f"(l2) : "f"(m));
    constexpr float ln2_hi = 0.693145751953125f;          // 16 significant bits: e * ln2_hi is exact
    constexpr float ln2_lo = 1.42860682030941723212e-6f;
    constexpr float ln2 = 0.69314718055994530942f;
    return fmaf(e, ln2_hi, fmaf(l2, ln2, e * ln2_lo));
}

// ---- mbarrier + 1-D bulk copy (TMA) ---------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Samples [lo, hi) of the clip that the bulk copy stages for the sub-tile starting at sample s0
// (stage[kStageLead + i] = x[s0 + i]).  Both ends are multiples of 4 samples so that the copy is
// 16-byte aligned and sized; everything else (the <= 3 tail samples of a clip, or the whole range when
// the clip is not 16-byte aligned in memory) is read with plain loads in the pre-pass.
struct StageRange { int lo, hi; };
__device__ __forceinline__ StageRange stage_range(int s0, int n, bool aligned) {
    StageRange r;
    r.lo = max(s0 - kStageLead, 0);
    r.hi = aligned ? min(s0 + kSubSamples, n & ~3) : r.lo;
    if (r.hi < r.lo) r.hi = r.lo;
    return r;
}

template <bool kPeak>
__global__ void __launch_bounds__(kThreads, 2)
k_frames(const float* __restrict__ pcm, const long long* __restrict__ offsets, const int* __restrict__ lengths,
         const float* __restrict__ peaks, const KTables* __restrict__ tab, int T_pad, int chunks_per_clip,
         float* __restrict__ out, double* __restrict__ partials) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const int b = blockIdx.y;
    const int chunk = blockIdx.x;
    const int n = lengths[b];
    const int T = n >= kFrame ? 1 + (n - kFrame) / kHop : 0;
    const int t_begin = chunk * kChunk;
    if (t_begin >= T) return;                       // uniform per CTA
    const int t_end = min(T, t_begin + kChunk);
    const float* clip = pcm + offsets[b];
    const bool aligned = (reinterpret_cast<unsigned long long>(clip) & 15ull) == 0;
    const float peak = kPeak ? peaks[b] : 1.0f;
    float* out_b = out + (size_t)b * T_pad * kMel;

    const int tid = threadIdx.x;
    const int slot = tid >> 4;
    const int r = tid & 15;
    const int lane = tid & 31;

    // first sub-tile's PCM: issue the bulk copy before anything else
    if (tid == 0) {
        mbar_init(&sm.mbar, 1);
        sm.stage[kStageLead - 1] = 0.0f;             // x[-1] of the clip's first sub-tile (later copies overwrite it)
        const StageRange sr = stage_range(t_begin * kHop, n, aligned);
        if (sr.hi > sr.lo) {
            mbar_expect_tx(&sm.mbar, (unsigned)(sr.hi - sr.lo) * 4u);
            bulk_g2s(sm.stage + (sr.lo - (t_begin * kHop - kStageLead)), clip + sr.lo, (unsigned)(sr.hi - sr.lo) * 4u, &sm.mbar);
        }
    }
    // tables -> shared
    for (int i = tid; i < kWinPad; i += kThreads) sm.win[i] = tab->win[i];
    sm.tw[tid] = tab->tw[tid];
    sm.post[tid] = tab->post[tid];
    for (int i = tid; i < kMelWeights; i += kThreads) sm.melw[i] = tab->melw[i];
    if (tid < kMel) sm.melmeta[tid] = tab->melmeta[tid];
    __syncthreads();                                // mbarrier init + tables visible

    // mel stage mapping: lane <-> frame, 16 groups of 5 mel bins (g, g + 16, ..., g + 64)
    const int mf = tid & 15, mg = tid >> 4;
    double s1[5], s2[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) s1[i] = s2[i] = 0.0;
    float* exf = reinterpret_cast<float*>(sm.ex);
    unsigned parity = 0;

    for (int t0 = t_begin; t0 < t_end; t0 += kSlots) {
        // ---- pre-pass: staged PCM -> d[i] = x[i] - 0.97 x[i-1] (float64) ----
        const int s0 = t0 * kHop;
        const StageRange sr = stage_range(s0, n, aligned);
        if (sr.hi > sr.lo) { mbar_wait(&sm.mbar, parity); parity ^= 1; }
        auto sample = [&](int g) -> float {          // x[g] of this clip, 0 outside
            float v = 0.0f;
            if (g >= sr.lo && g < sr.hi) v = sm.stage[g - s0 + kStageLead];
            else if (g >= 0 && g < n) v = __ldg(clip + g);
            if (kPeak) v = v / peak;                  // float32 division, like numpy's (R/processor.py:92)
            return v;
        };
        const bool full = sr.hi == s0 + kSubSamples && (s0 == 0 || sr.lo == s0 - kStageLead);
        if (full) {
            // whole sub-tile staged and inside the clip: 4 samples per thread and step, no range checks
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int q = tid + u * kThreads;
                if (q < kSubSamples / 4) {
                    float4 x4 = *reinterpret_cast<const float4*>(sm.stage + kStageLead + 4 * q);
                    float xp = sm.stage[kStageLead - 1 + 4 * q];
                    if (kPeak) { x4.x /= peak; x4.y /= peak; x4.z /= peak; x4.w /= peak; xp /= peak; }
                    const double x0 = (double)x4.x, x1 = (double)x4.y, x2 = (double)x4.z, x3 = (double)x4.w;
                    double2* d2 = reinterpret_cast<double2*>(sm.dtile + 4 * q);
                    d2[0] = make_double2(fma(-0.97, (double)xp, x0), fma(-0.97, x0, x1));
                    d2[1] = make_double2(fma(-0.97, x1, x2), fma(-0.97, x2, x3));
                }
            }
        } else {
#pragma unroll 1
            for (int u = 0; u < kSubSamples / kThreads; ++u) {
                const int i = tid + u * kThreads;
                sm.dtile[i] = fma(-0.97, (double)sample(s0 + i - 1), (double)sample(s0 + i));
            }
        }
        if (tid < kSlots) {
            sm.xb0[tid] = (double)sample(s0 + tid * kHop);
            sm.xb1[tid] = (double)sample(s0 + tid * kHop + kFrame - 1);
        }
        __syncthreads();                            // dtile ready; staging and the exchange area are free again

        // next sub-tile's PCM lands while this one is transformed
        if (tid == 0 && t0 + kSlots < t_end) {
            const StageRange nx = stage_range(s0 + kSlots * kHop, n, aligned);
            if (nx.hi > nx.lo) {
                mbar_expect_tx(&sm.mbar, (unsigned)(nx.hi - nx.lo) * 4u);
                bulk_g2s(sm.stage + (nx.lo - (s0 + kSlots * kHop - kStageLead)), clip + nx.lo,
                         (unsigned)(nx.hi - nx.lo) * 4u, &sm.mbar);
            }
        }

        // ---- frame -> y (float64) ----
        const double2* dfr = reinterpret_cast<const double2*>(sm.dtile + slot * kHop);
        const double2* wfr = reinterpret_cast<const double2*>(sm.win);
        cd v[16];
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            double2 p = dfr[r + 16 * j];
            v[j] = {p.x, p.y};
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) s += v[j].re + v[j].im;
        if (r == 0) s -= v[0].re;                    // i = 0 is not part of sum_{i=1..399} d[i]
        if (r < 8) s += v[12].re + v[12].im;         // i = 2 (r + 192) (+1) <= 399  <=>  r <= 7
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        // 0.03 * sum(x) = sum_{i>=1} d[i] + x[0] - 0.97 x[399]
        const double c = (s + sm.xb0[slot] - 0.97 * sm.xb1[slot]) * (1.0 / 400.0);
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            double2 w = wfr[r + 16 * j];
            v[j].re = w.x * (v[j].re - c);
            v[j].im = w.y * (v[j].im - c);
        }

        // ---- pass 1 ----  exchange element (row k1, column n2) lives at k1 * 16 + (n2 ^ (k1 & 7))
        cd a[16];
        dft16<true>(v, a);
        double2* ex = sm.ex + slot * 256;
        ex[r] = make_double2(a[0].re, a[0].im);
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) {
            double2 w = sm.tw[k1 * 16 + r];
            cd m = cmul(a[k1], w.x, w.y);
            ex[k1 * 16 + (r ^ (k1 & 7))] = make_double2(m.re, m.im);
        }
        __syncwarp();

        // ---- pass 2 (thread r now owns row k1 = r) ----
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) {
            double2 p = ex[r * 16 + (n2 ^ (r & 7))];
            v[n2] = {p.x, p.y};
        }
        dft16<false>(v, a);                          // a[k2] = Z[r + 16 k2]
        __syncwarp();                                // the slot's exchange area is now free for its power spectrum

        // ---- real split + power ----
        float* P = exf + slot * kSlotFloats + p_skew(slot);      // 256 floats inside the slot's own area
        const int partner = (lane & 16) | ((16 - r) & 15);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
            double pr = __shfl_sync(0xffffffffu, a[15 - k2].re, partner);
            double pi = __shfl_sync(0xffffffffu, a[15 - k2].im, partner);
            pr = r == 0 ? a[(16 - k2) & 15].re : pr;             // thread 0 pairs Z[16 k2] with its own Z[256 - 16 k2]
            pi = r == 0 ? a[(16 - k2) & 15].im : pi;
            const double2 w = sm.post[k2 * 16 + r];
            const double ar = a[k2].re, ai = a[k2].im;
            const double sr_ = ar + pr, dr = ar - pr, si = ai + pi, di = ai - pi;
            const double xr = fma(w.y, dr, fma(w.x, si, sr_));      // 2 Re X[k]
            const double xi = fma(w.y, si, fma(-w.x, dr, di));      // 2 Im X[k]
            const float fr = (float)xr, fi = (float)xi;             // the reference rounds X to complex64
            P[r + 16 * k2] = fmaf(fr, fr, fi * fi);                 // 4 |X|^2 (the 1/4 is in the mel weights)
        }
        __syncthreads();                            // all 16 power spectra visible

        // ---- sparse mel + ln: lane <-> frame (weights broadcast, power spectra conflict-free) ----
        {
            const float* Pf = exf + mf * kSlotFloats + p_skew(mf);
            float* orow = exf + 512 + mf * kOutRow;
            const double on = (t0 + mf < t_end) ? 1.0 : 0.0;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int m = mg + 16 * i;
                const int meta = sm.melmeta[m];
                const int first = meta & 511, count4 = (meta >> 9) & 127, off4 = meta >> 16;
                const float4* w4 = reinterpret_cast<const float4*>(sm.melw) + off4;
                const float* pk = Pf + first;
                float acc = 0.0f;
                for (int q = 0; q < count4; ++q) {
                    const float4 w = w4[q];
                    acc = fmaf(w.x, pk[4 * q + 0], acc);
                    acc = fmaf(w.y, pk[4 * q + 1], acc);
                    acc = fmaf(w.z, pk[4 * q + 2], acc);
                    acc = fmaf(w.w, pk[4 * q + 3], acc);
                }
                const float lg = ln_pos(fmaxf(acc, kMelFloor));
                orow[m] = lg;
                const double lgd = (double)lg;
                s1[i] = fma(on, lgd, s1[i]);
                s2[i] = fma(on * lgd, lgd, s2[i]);
            }
        }
        __syncthreads();                            // staged rows complete

        // ---- coalesced store of the sub-tile's rows: frames t0 .. are one contiguous block of out ----
        {
            const int rows = min(min(t_end, T_pad) - t0, kSlots);
            float* dst = out_b + (size_t)t0 * kMel;
            for (int e = tid; e < rows * (kMel / 4); e += kThreads) {
                const int f = e / (kMel / 4), m4 = (e - f * (kMel / 4)) * 4;
                const float* src = exf + 512 + f * kOutRow + m4;
                reinterpret_cast<float4*>(dst)[e] = make_float4(src[0], src[1], src[2], src[3]);
            }
        }
        // the next iteration's pre-pass only touches dtile / xb / stage; its __syncthreads orders these
        // reads of the exchange area before the next pass 1 overwrites it
    }

    // ---- per-chunk statistics: ordered reduction over the 16 frame lanes ----
    __syncthreads();
    constexpr int kRedRow = kMel + 1;                  // odd stride: the 16 frame lanes hit distinct banks
    double* red = reinterpret_cast<double*>(sm.ex);   // [2][16][81]
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        red[mf * kRedRow + mg + 16 * i] = s1[i];
        red[kSlots * kRedRow + mf * kRedRow + mg + 16 * i] = s2[i];
    }
    __syncthreads();
    if (tid < 2 * kMel) {
        const int which = tid / kMel, m = tid - which * kMel;
        double acc = 0.0;
#pragma unroll
        for (int sl = 0; sl < kSlots; ++sl) acc += red[which * kSlots * kRedRow + sl * kRedRow + m];
        partials[((size_t)b * chunks_per_clip + chunk) * (2 * kMel) + tid] = acc;
    }
}

// mean and 1/sqrt(var + 1e-7) per (clip, bin); var with ddof = 1 (…seamless_m4t.py:257-262)
__global__ void k_finalize(const int* __restrict__ lengths, const double* __restrict__ partials,
                           int chunks_per_clip, double* __restrict__ stats) {
    const int b = blockIdx.x, m = threadIdx.x;
    if (m >= kMel) return;
    const int n = lengths[b];
    const int T = n >= kFrame ? 1 + (n - kFrame) / kHop : 0;
    const int nchunks = (T + kChunk - 1) / kChunk;
    double a1 = 0.0, a2 = 0.0;
    for (int c = 0; c < nchunks; ++c) {
        const double* p = partials + ((size_t)b * chunks_per_clip + c) * (2 * kMel);
        a1 += p[m];
        a2 += p[kMel + m];
    }
    const double mean = a1 / (double)T;
    // a single frame has no ddof=1 variance: numpy returns NaN there and so do we
    double var = T > 1 ? (a2 - a1 * mean) / (double)(T - 1) : __longlong_as_double(0x7ff8000000000000LL);
    if (var < 0.0) var = 0.0;                                      // rounding of a constant column; keeps NaN
    stats[((size_t)b * kMel + m) * 2 + 0] = mean;
    stats[((size_t)b * kMel + m) * 2 + 1] = 1.0 / sqrt(var + 1e-7);
}

// in-place CMVN + padding rows + mask.  One thread per float4 of a clip's [T_pad, 80] block.
__global__ void __launch_bounds__(256)
k_normalize(const int* __restrict__ lengths, const double* __restrict__ stats, int T_pad, float padding_value,
            int normalize, float* __restrict__ out, int* __restrict__ mask) {
    __shared__ double s_mean[kMel], s_rstd[kMel];
    const int b = blockIdx.y;
    const int n = lengths[b];
    const int T = min(n >= kFrame ? 1 + (n - kFrame) / kHop : 0, T_pad);
    if (threadIdx.x < kMel) {
        s_mean[threadIdx.x] = stats[((size_t)b * kMel + threadIdx.x) * 2 + 0];
        s_rstd[threadIdx.x] = stats[((size_t)b * kMel + threadIdx.x) * 2 + 1];
    }
    __syncthreads();
    const int quads = T_pad * (kMel / 4);
    float4* o4 = reinterpret_cast<float4*>(out + (size_t)b * T_pad * kMel);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
        const int t = q / (kMel / 4), m = (q - t * (kMel / 4)) * 4;
        float4 v;
        if (t < T) {
            if (!normalize) continue;
            v = o4[q];
            v.x = (float)(((double)v.x - s_mean[m + 0]) * s_rstd[m + 0]);
            v.y = (float)(((double)v.y - s_mean[m + 1]) * s_rstd[m + 1]);
            v.z = (float)(((double)v.z - s_mean[m + 2]) * s_rstd[m + 2]);
            v.w = (float)(((double)v.w - s_mean[m + 3]) * s_rstd[m + 3]);
        } else {
            v = make_float4(padding_value, padding_value, padding_value, padding_value);
        }
        o4[q] = v;
    }
    if (mask) {
        const int rows = T_pad / 2;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < rows; j += gridDim.x * blockDim.x)
            mask[(size_t)b * rows + j] = (2 * j + 1 < T) ? 1 : 0;
    }
}

__global__ void k_peak_init(float* __restrict__ peaks, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) peaks[i] = 1.0f;
}

// max(1, max|x|) per clip; non-negative floats order like their bit patterns
__global__ void __launch_bounds__(256)
k_peak_abs(const float* __restrict__ pcm, const long long* __restrict__ offsets, const int* __restrict__ lengths,
           float* __restrict__ peaks) {
    const int b = blockIdx.y;
    const int n = lengths[b];
    const float* clip = pcm + offsets[b];
    float m = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(__ldg(clip + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 1.0f) atomicMax(reinterpret_cast<int*>(peaks + b), __float_as_int(m));
}

// ---- device tables, one copy per device --------------------------------------------------
std::mutex g_tab_mutex;
KTables* g_tab[64] = {nullptr};

int get_tables(const KTables** out) {
    int dev = 0;
    STX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return STX_EINVAL; }
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    if (!g_tab[dev]) {
        static KTables h;   // ~13 KB, filled once
        const std::vector<double>& w = k_window();
        for (int i = 0; i < kWinPad; ++i) h.win[i] = i < kFrame ? w[i] * 32768.0 : 0.0;
        for (int k1 = 0; k1 < 16; ++k1)
            for (int r = 0; r < 16; ++r) {
                double ang = -2.0 * M_PI * double(r * k1) / 256.0;
                h.tw[k1 * 16 + r] = make_double2(std::cos(ang), std::sin(ang));
            }
        for (int k2 = 0; k2 < 16; ++k2)
            for (int k1 = 0; k1 < 16; ++k1) {
                double ang = -2.0 * M_PI * double(k1 + 16 * k2) / 512.0;
                h.post[k2 * 16 + k1] = make_double2(std::cos(ang), std::sin(ang));
            }
        MelCsr csr = build_mel_csr(k_mel(), STX_K_NFFT / 2 + 1, kMel, 0.25, 4, 256);
        if (csr.weights.size() > size_t(kMelWeights)) { set_error("mel table overflow"); return STX_EINVAL; }
        for (int i = 0; i < kMelWeights; ++i) h.melw[i] = i < int(csr.weights.size()) ? csr.weights[i] : 0.0f;
        for (int m = 0; m < kMel; ++m) {
            if (csr.first[m] + csr.count[m] > 256) { set_error("mel filter %d reaches the Nyquist bin", m); return STX_EINVAL; }
            h.melmeta[m] = csr.first[m] | ((csr.count[m] / 4) << 9) | ((csr.offset[m] / 4) << 16);
        }
        KTables* d = nullptr;
        STX_CUDA(cudaMalloc(&d, sizeof(KTables)));
        STX_CUDA(cudaMemcpy(d, &h, sizeof(KTables), cudaMemcpyHostToDevice));
        STX_CUDA(cudaFuncSetAttribute(k_frames<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        STX_CUDA(cudaFuncSetAttribute(k_frames<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        g_tab[dev] = d;
    }
    *out = g_tab[dev];
    return 0;
}

inline int frames_of(int n) { return n >= kFrame ? 1 + (n - kFrame) / kHop : 0; }
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace
}  // namespace stx

extern "C" {

int stx_fbank_k_workspace(int B, int max_length, size_t* bytes) {
    using namespace stx;
    if (B < 0 || max_length < 0 || !bytes) { set_error("stx_fbank_k_workspace: bad argument"); return STX_EINVAL; }
    const int chunks = (frames_of(max_length) + kChunk - 1) / kChunk;
    *bytes = align256(size_t(B) * std::max(chunks, 1) * 2 * kMel * sizeof(double)) +
             align256(size_t(B) * kMel * 2 * sizeof(double));
    return 0;
}

int stx_fbank_k(const float* d_pcm, const int64_t* d_offsets, const int32_t* d_lengths, int B, int max_length,
                const float* d_peak, int T_pad, float padding_value, int normalize, float* d_out,
                int32_t* d_mask, void* d_ws, size_t ws_bytes, void* stream) {
    using namespace stx;
    if (B < 0 || max_length < 0 || T_pad < 0 || (T_pad & 1)) { set_error("stx_fbank_k: B, max_length >= 0 and even T_pad required"); return STX_EINVAL; }
    if (B == 0 || T_pad == 0) return 0;
    if (!d_pcm || !d_offsets || !d_lengths || !d_out || !d_ws) { set_error("stx_fbank_k: null pointer"); return STX_EINVAL; }
    if (int rc = check_device()) return rc;
    size_t need = 0;
    stx_fbank_k_workspace(B, max_length, &need);
    if (ws_bytes < need) { set_error("stx_fbank_k: workspace %zu < %zu bytes", ws_bytes, need); return STX_ENOSPACE; }
    if (B > 65535) { set_error("stx_fbank_k: B = %d > 65535 clips per call", B); return STX_EINVAL; }
    const KTables* tab = nullptr;
    if (int rc = get_tables(&tab)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    const int chunks = std::max((frames_of(max_length) + kChunk - 1) / kChunk, 1);
    double* partials = static_cast<double*>(d_ws);
    double* stats = reinterpret_cast<double*>(static_cast<char*>(d_ws) +
                                              align256(size_t(B) * chunks * 2 * kMel * sizeof(double)));
    if (frames_of(max_length) > 0) {
        if (d_peak) {
            STX_LAUNCH(k_frames<true>, dim3(chunks, B), dim3(kThreads), sizeof(Smem), st,
                       d_pcm, reinterpret_cast<const long long*>(d_offsets), d_lengths, d_peak, tab, T_pad, chunks,
                       d_out, partials);
        } else {
            STX_LAUNCH(k_frames<false>, dim3(chunks, B), dim3(kThreads), sizeof(Smem), st,
                       d_pcm, reinterpret_cast<const long long*>(d_offsets), d_lengths, d_peak, tab, T_pad, chunks,
                       d_out, partials);
        }
    }
    STX_LAUNCH(k_finalize, dim3(B), dim3(96), 0, st, d_lengths, partials, chunks, stats);
    const int quads = T_pad * (kMel / 4);
    const int gx = std::max(1, std::min((quads + 255) / 256, 64));
    STX_LAUNCH(k_normalize, dim3(gx, B), dim3(256), 0, st, d_lengths, stats, T_pad, padding_value, normalize,
               d_out, d_mask);
    return 0;
}

int stx_peak_abs(const float* d_pcm, const int64_t* d_offsets, const int32_t* d_lengths, int B, float* d_peak,
                 void* stream) {
    using namespace stx;
    if (B < 0) { set_error("stx_peak_abs: B < 0"); return STX_EINVAL; }
    if (B == 0) return 0;
    if (!d_pcm || !d_offsets || !d_lengths || !d_peak) { set_error("stx_peak_abs: null pointer"); return STX_EINVAL; }
    if (B > 65535) { set_error("stx_peak_abs: B = %d > 65535", B); return STX_EINVAL; }
    if (int rc = check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    STX_LAUNCH(k_peak_init, dim3((B + 255) / 256), dim3(256), 0, st, d_peak, B);
    STX_LAUNCH(k_peak_abs, dim3(32, B), dim3(256), 0, st, d_pcm, reinterpret_cast<const long long*>(d_offsets),
               d_lengths, d_peak);
    return 0;
}

}  // extern "C"
