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
//   w_frames    one CTA per (clip, chunk of frames), ONE FRAME PER LANE like k_frames (fbank_k.cu): a tile is 32
//               consecutive frames and the 8 warps split each frame's 400-point real FFT into 16 roles, two per
//               warp in sequence (n = 16 n1 + n2, k = k1 + 25 k2).  In float32 the exchange is 50 KB per tile, so
//               two CTAs fit on an SM and overlap each other's phases:
//                 stage    cp.async.bulk (TMA) of the tile's raw samples, one tile ahead
//                 layout   reflect / zero padding and the peak divisor applied once per sample, rows of 161 floats
//                 pass 1   warp n2: Hann window, real DFT-25 over n1 in registers (codelets.cuh), times W400^(n2 k1)
//                 pass 2   warp k1 = 1..12: complex DFT-16 over n2 -> X[k1 + 25 k2] (k >= 201 are the mirrored bins);
//                          warp 0: the real row k1 = 0
//                 power -> shared [bin][lane]; mel (fixed-length padded filters) -> log10 -> coalesced stores
//                 along t, per-clip max by warp reduce + integer atomics
//   w_finish    in-place max(x, max - 8), (x + 4) / 4 and the mask
//
// DFT engine: a shared-memory / register FFT on the FP32 pipe.  The alternative the task names, DFT-as-GEMM on
// tcgen05 with split operands, needs 2 * 400 * 402 flop per frame and pass (x3 passes for float32-grade accuracy:
// 185 GFLOP for the 192 000 frames of the 64 x 30 s batch, >= 130 us at the measured 1.4 PFLOP/s), against
// ~7.3 kflop per frame for the FFT; see DESIGN.md §5.
#include "stx_common.h"
#include "codelets.cuh"
#include <cmath>
#include <cstdlib>
#include <cstddef>
#include <mutex>

namespace stx {
namespace {

constexpr int kN = STX_W_NFFT;          // 400
constexpr int kHop = STX_W_HOP;         // 160
constexpr int kMel = STX_W_NMEL;        // 80
constexpr int kBins = kN / 2 + 1;       // 201
constexpr int kTile = 32;               // frames per tile = lanes
constexpr int kWarps = 8;                // every warp takes two of the 16 roles, one after the other
constexpr int kThreads = kWarps * 32;   // 256 threads, ~99 KB of shared memory: two CTAs per SM, so that one CTA's
                                        // shared-memory phases overlap the other's arithmetic
constexpr int kTileSamples = (kTile - 1) * kHop + kN;   // 5360
constexpr int kXRow = kHop + 1;         // padded rows: odd stride, lane f reads row f + const without bank conflicts
constexpr int kXBuf = 34 * kXRow;
__host__ __device__ constexpr int mel_len(int slot) { return slot < 3 ? 4 : slot == 3 ? 8 : 16; }
__host__ __device__ constexpr int mel_off(int slot) { return slot == 0 ? 0 : mel_off(slot - 1) + 16 * mel_len(slot - 1); }
constexpr int kMelWeights = mel_off(5);  // 576
constexpr int kStockSamples = 480000;   // the recipe's stock chunk: 30 s (w_frames is specialised for it)
#ifndef STX_W_MEL_EXACT
#define STX_W_MEL_EXACT 1               // 1 (shipped): the mel stage walks ten half slots of 8 filters (one per warp) whose lengths are
                                        //    the longest filter of the half slot (2, 2, 2, 2, 4, 4, 6, 8, 11, 14 bins: 440 bin reads and
                                        //    FMAs per frame) instead of five slots of 16 padded to 4, 4, 4, 8, 16 (576); the dropped
                                        //    products had zero weights, so the results are bit-identical
#endif
// half slot j holds mel bins 8 j .. 8 j + 7 (one per warp): mel_len10 bins are read, the weights are stored padded to a multiple
// of four (aligned float4 loads)
__host__ __device__ constexpr int mel_len10(int j) {
    return j < 4 ? 2 : j < 6 ? 4 : j == 6 ? 6 : j == 7 ? 8 : j == 8 ? 11 : 14;
}
__host__ __device__ constexpr int mel_pad10(int j) { return (mel_len10(j) + 3) & ~3; }
__host__ __device__ constexpr int mel_off10(int j) { return j == 0 ? 0 : mel_off10(j - 1) + 8 * mel_pad10(j - 1); }
static_assert(mel_off10(10) <= kMelWeights, "the half-slot table fits the array of the slot table");

__constant__ __align__(8) float cw_win[16][26];     // [n2][n1] = hann[16 n1 + n2]; rows padded to 26: read as 13 aligned pairs (LDCU.64)
__constant__ float2 cw_tw[16][16];      // [n2][k1] = W400^(n2 k1), k1 = 0..12

struct WTables {
    float melw[kMelWeights];            // [slot][warp][mel_len(slot)]
    int   melfirst[kMel];
};

// One 32-frame tile of a work item (clip b, chunk of frames), kept in shared memory and read where it is used
struct WTile {
    const float* clip; float* out_b; unsigned* cmax;
    float* fmin_b;                      // (STX_W_LAZY_CLAMP) [8 warps][T]: smallest log10 of the warp's ten mel bins per frame
    int len, t0, t_end, item;
    float peak;
    int aligned, valid, last;           // last: the item ends with this tile (publish the running maximum)
};
struct Smem {
    float2 ex[12][16][kTile];           // pass-1 rows k1 = 1..12: [k1 - 1][n2][lane]
    float  ex0[16][kTile];              // row k1 = 0 (real)
    union {
        float xs[kXBuf + 2];            // padded, windowable samples of the tile (dead after pass 1)
        float P[kBins][kTile];          // power spectrum [bin][lane] (pass 2 -> mel)
    } u;
    float  stage[kTileSamples];         // raw PCM of the next tile (cp.async.bulk)
    float  melw[kMelWeights];
    int    melfirst[kMel];
    WTile  desc[2];                     // the tile in flight and the next one (written by thread 0)
    unsigned long long mbar;
};
static_assert(sizeof(Smem) + 1024 <= 233472 / 2, "two CTAs per SM: 2 x (dynamic + 1 KB reserved) <= 228 KB");
static_assert(offsetof(Smem, stage) % 16 == 0, "bulk-copy destination alignment");

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
// evict-first in L2: the PCM is read once, the raw log-mel written by this kernel is re-read by w_finish
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// Clip samples [lo, hi) that the bulk copy stages for the tile whose first padded sample is g0 = 160 t0 - 200
// (stage[g - g0] = x[g]); multiples of 4 samples; the rest (reflected / tail samples, unaligned clips) are plain loads.
struct StageRange { int lo, hi; };
__device__ __forceinline__ StageRange stage_range(int g0, int len, bool aligned) {
    StageRange r;
    r.lo = max(g0, 0);
    r.hi = aligned ? min(g0 + kTileSamples, len & ~3) : r.lo;
    if (r.hi < r.lo) r.hi = r.lo;
    return r;
}

// log10(x) for normal positive x (the 1e-10 floor guarantees it): exponent + MUFU.LG2 of the mantissa, times log10(2).
// The mantissa's log2 is in [0, 1), where lg2.approx is accurate to 2^-22 absolute, so the result is within 1e-7 of
// log10f at a third of its instructions (the bar on log10 is 4e-4).
#ifndef STX_W_LAZY_CLAMP
#define STX_W_LAZY_CLAMP 1              // 1 (shipped): w_frames writes the final (x + 4) / 4 itself and leaves, per frame and warp, the
                                        //    smallest log10 it produced; w_finish then touches only the frames that the per-clip clamp
                                        //    max(x, max - 8) actually changes (max commutes with the monotone (x + 4) / 4, so the bits are
                                        //    those of clamp-then-scale).  For audio without 8 decades of dynamic range inside a clip --
                                        //    anything but digital silence and zero padding -- the second pass over the 61 MB of cfg2
                                        //    becomes a 6 MB read.  0: w_finish rewrites every value (rounds 1 and 2)
#endif
static_assert(!STX_W_LAZY_CLAMP || STX_W_MEL_EXACT, "the lazy clamp is written into the half-slot mel stage");
#ifndef STX_W_LOG_FAST
#define STX_W_LOG_FAST 1                // 1 (shipped): log10 of a mel energy as lg2.approx of the whole value times log10(2): 2 instructions
                                        //    instead of 8.  The result is a float32 of magnitude up to 33 in log2, i.e. good to 3.8e-6 there,
                                        //    1.1e-6 in log10 and 3e-7 on the (x + 4) / 4 the bar of 1e-4 applies to
#endif
__device__ __forceinline__ float log10_pos(float x) {
#if STX_W_LOG_FAST
    float l2x;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2x) : "f"(x));
    return l2x * 0.30102999566398120f;
#endif
    const int bits = __float_as_int(x);
    const float e = (float)((bits >> 23) - 127);
    const float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(m));     // m is in [1, 2): .ftz only drops the denormal-input fix-up code
    constexpr float k_hi = 0.30102539062500f;             // log10(2) to 12 significant bits: e * k_hi is exact
    constexpr float k_lo = 4.6050389811952137e-6f;        // log10(2) - k_hi
    constexpr float k = 0.30102999566398120f;
    return fmaf(e, k_hi, fmaf(l2, k, e * k_lo));
}

template <int kSlot>
__device__ __forceinline__ float mel_slot(const float* __restrict__ Pl, const float* __restrict__ melw,
                                          const int* __restrict__ melfirst, int warp) {
    constexpr int L = mel_len(kSlot);
    const float4* w4 = reinterpret_cast<const float4*>(melw + mel_off(kSlot) + warp * L);
    const float* pk = Pl + melfirst[16 * kSlot + warp] * kTile;
    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
    for (int q = 0; q < L / 4; ++q) {
        const float4 w = w4[q];
        acc0 = fmaf(w.x, pk[(4 * q + 0) * kTile], acc0);
        acc1 = fmaf(w.y, pk[(4 * q + 1) * kTile], acc1);
        acc0 = fmaf(w.z, pk[(4 * q + 2) * kTile], acc0);
        acc1 = fmaf(w.w, pk[(4 * q + 3) * kTile], acc1);
    }
    return log10_pos(fmaxf(acc0 + acc1, 1e-10f));
}

template <int kHalfSlot>
__device__ __forceinline__ float mel_slot10(const float* __restrict__ Pl, const float* __restrict__ melw,
                                            const int* __restrict__ melfirst, int warp) {
    constexpr int L = mel_len10(kHalfSlot), PAD = mel_pad10(kHalfSlot);
    const float4* w4 = reinterpret_cast<const float4*>(melw + mel_off10(kHalfSlot) + warp * PAD);
    const float* pk = Pl + melfirst[8 * kHalfSlot + warp] * kTile;
    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
    for (int q = 0; q < PAD / 4; ++q) {
        const float4 w = w4[q];
        if (4 * q + 0 < L) acc0 = fmaf(w.x, pk[(4 * q + 0) * kTile], acc0);
        if (4 * q + 1 < L) acc1 = fmaf(w.y, pk[(4 * q + 1) * kTile], acc1);
        if (4 * q + 2 < L) acc0 = fmaf(w.z, pk[(4 * q + 2) * kTile], acc0);
        if (4 * q + 3 < L) acc1 = fmaf(w.w, pk[(4 * q + 3) * kTile], acc1);
    }
    return log10_pos(fmaxf(acc0 + acc1, 1e-10f));
}

// Per-clip max through an order-preserving integer key (negative floats: all bits flipped, others: sign bit set), so
// that key 0 is below every float and a plain memset initialises the maxima (no init kernel).
__device__ __forceinline__ unsigned max_key(float v) {
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float max_unkey(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// PERSISTENT: two CTAs per SM, every CTA streams through the work items (clip b, chunk of chunk_frames frames) blockIdx.x,
// blockIdx.x + gridDim.x, ...: tables, barrier initialisation and the pipeline fill are paid once per kernel instead of
// once per chunk (1.5 us each, 9 % of the non-persistent form), and the PCM of an item's first tile lands while the
// previous item's last tile is transformed.
// Launch bound of THREE CTAs per SM although shared memory admits two: at the 80 registers that bound implies ptxas reads the
// per-role window / twiddle tables through the uniform datapath (84 LDCU, no spills); at 126 registers it reads them with
// register-indexed LDCs, which queue in the MIO behind the shared-memory traffic (the two hottest lines of the round-1 kernel).
// kFixedSamples: n_samples at compile time (the stock 30 s chunk: 480 000), 0 = taken from the argument.  With T = n_samples / 160
// a constant, the ten output rows of a warp ([B, 80, T]: rows are T floats apart) are immediate offsets from one pointer
// instead of a 64-bit multiply-add and a two-instruction address per store (40 of ~330 instructions of the mel stage).
template <int kFixedSamples>
__global__ void __launch_bounds__(kThreads, 3)
w_frames(const float* __restrict__ pcm, const long long* __restrict__ offsets, const int* __restrict__ lengths,
         const float* __restrict__ peaks, const WTables* __restrict__ tab, int B, int n_samples_arg, int chunk_frames,
         int chunks_per_clip, float* __restrict__ out, unsigned* __restrict__ clip_max, float* __restrict__ frame_min) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const int n_samples = kFixedSamples ? kFixedSamples : n_samples_arg;
    const int T = n_samples / kHop;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int items = B * chunks_per_clip;

    // (thread 0 only) first tile of the item; every item is non-empty (all clips are padded to n_samples)
    auto open_item = [&](int item, WTile& d) {
        d.valid = 0; d.last = 0; d.item = item;
        if (item < items) {
            const int b = item / chunks_per_clip, chunk = item - b * chunks_per_clip;
            d.valid = 1; d.t0 = chunk * chunk_frames; d.t_end = min(T, d.t0 + chunk_frames);
            d.len = min(__ldg(lengths + b), n_samples);
            d.clip = pcm + __ldg(offsets + b);
            d.aligned = (reinterpret_cast<unsigned long long>(d.clip) & 15ull) == 0;
            d.peak = peaks ? __ldg(peaks + b) : 1.0f;
            d.out_b = out + (size_t)b * kMel * T;
            d.cmax = clip_max + b;
            d.fmin_b = frame_min + (size_t)b * kWarps * T;
        }
    };
    auto next_tile = [&](WTile& c, WTile& d) {       // (thread 0 only) d = the tile after c
        if (c.t0 + kTile < c.t_end) { d = c; d.t0 += kTile; c.last = 0; }
        else { open_item(c.item + (int)gridDim.x, d); c.last = 1; }
    };
    auto prefetch = [&](const WTile& d) {           // (thread 0 only) the tile's PCM -> staging; nothing for padding-only tiles
        if (d.valid) {
            const int g0 = d.t0 * kHop - kN / 2;
            const StageRange sr = stage_range(g0, d.len, d.aligned != 0);
            if (sr.hi > sr.lo) {
                mbar_expect_tx(&sm.mbar, (unsigned)(sr.hi - sr.lo) * 4u);
                bulk_g2s(sm.stage + (sr.lo - g0), d.clip + sr.lo, (unsigned)(sr.hi - sr.lo) * 4u, &sm.mbar);
            }
        }
    };
    if (tid == 0) {
        mbar_init(&sm.mbar, 1);
        open_item(blockIdx.x, sm.desc[0]);
        prefetch(sm.desc[0]);
    }
    for (int i = tid; i < kMelWeights; i += kThreads) sm.melw[i] = tab->melw[i];
    if (tid < kMel) sm.melfirst[tid] = tab->melfirst[tid];
    __syncthreads();

    float run_max = __int_as_float(0xff800000);
    unsigned parity = 0;
    int slot = 0;

    while (sm.desc[slot].valid) {
        const WTile& cur = sm.desc[slot];
        const int t0 = cur.t0, t_end = cur.t_end, len = cur.len;
        const float* clip = cur.clip;
        const bool aligned = cur.aligned != 0;
        const float peak = cur.peak;
        float* out_b = cur.out_b;
        // ---- layout: zero padding to n_samples, reflect padding of 200 around it, peak divisor; rows of 161 ----
        const int g0 = t0 * kHop - kN / 2;
        const bool padding_only = g0 >= len && len <= n_samples - (kN / 2 + 2);
        if (padding_only) {
            // the whole tile lies in the zero padding behind the clip (and the reflected tail of the padded signal is
            // zero as well): every mel energy is 0, so every value is log10 of the floor; no copy was issued for it
            const float v = log10_pos(1e-10f);
            const int t = t0 + lane;
            if (t < t_end) {
#if STX_W_LAZY_CLAMP
                // nothing is written for the frame: the minimum of warp 0 carries a marker (-inf) and w_finish, which knows the
                // clip's floor by then, stores the final value of all 80 bins once (instead of a store here and a read-modify-write
                // there: the zero padding of a short clip is always below the floor unless the whole clip is silent)
                cur.fmin_b[(size_t)warp * T + t] = warp == 0 ? __int_as_float(0xff800000) : v;
#else
#pragma unroll
                for (int i = 0; i < 10; ++i) out_b[(size_t)(warp + 8 * i) * T + t] = v;
#endif
                run_max = fmaxf(run_max, v);
            }
            __syncthreads();                        // every warp is done with the previous tile's descriptor (the other slot)
            if (tid == 0) {
                next_tile(sm.desc[slot], sm.desc[slot ^ 1]);
                prefetch(sm.desc[slot ^ 1]);
            }
            __syncthreads();                        // the next descriptor and `last` are visible
        } else {
        const StageRange sr = stage_range(g0, len, aligned);
        if (sr.hi > sr.lo) { mbar_wait(&sm.mbar, parity); parity ^= 1; }
        if (sr.lo == g0 && sr.hi == g0 + kTileSamples && peak == 1.0f) {
            // all 256 threads, 21 elements each, every load issued before the first store (a store in between would make
            // the copies wait for each other: the pass is pure latency)
            float v[21];
#pragma unroll
            for (int j = 0; j < 21; ++j) {
                const int i = tid + kThreads * j;
                if (j < 20 || i < kTileSamples) v[j] = sm.stage[i];
            }
#pragma unroll
            for (int j = 0; j < 21; ++j) {
                const int i = tid + kThreads * j;
                if (j < 20 || i < kTileSamples) sm.u.xs[i + (unsigned)i / kHop] = v[j];
            }
        } else if (sr.lo == g0 && sr.hi == g0 + kTileSamples) {
#pragma unroll 1
            for (int i = tid; i < kTileSamples; i += kThreads)
                sm.u.xs[i + (unsigned)i / kHop] = sm.stage[i] / peak;    // float32 division, like numpy's (R/processor.py:92)
        } else {
#pragma unroll 1
            for (int i = tid; i < kTileSamples; i += kThreads) {
                int g = g0 + i;
                if (g < 0) g = -g;
                if (g >= n_samples) g = 2 * (n_samples - 1) - g;
                float x = 0.0f;
                if (g >= sr.lo && g < sr.hi) x = sm.stage[g - g0];
                else if (g >= 0 && g < len) x = __ldg(clip + g);
                if (peak != 1.0f) x = x / peak;
                sm.u.xs[i + (unsigned)i / kHop] = x;
            }
        }
        __syncthreads();                            // xs ready; staging is free

        // ---- window + pass 1 (role = n2 = warp, warp + 8) ----
        // two roles per warp, written out (not a loop): `warp` is provably warp-uniform, a loop-carried role is not, and only
        // a provably uniform index lets the window / twiddle reads go through the uniform datapath instead of queueing
        // register-indexed constant loads in the MIO with the shared-memory traffic
        auto pass1 = [&](const int role) {
            const float* X = sm.u.xs + kXRow * lane + role;
            float y[25], re[13], im[13];
#pragma unroll
            for (int n1 = 0; n1 < 25; n1 += 2) {       // (window values in pairs: half the uniform loads)
                const float2 wv = *reinterpret_cast<const float2*>(&cw_win[role][n1]);
                y[n1] = wv.x * X[16 * n1 + (n1 >= 10) + (n1 >= 20)];
                if (n1 + 1 < 25) y[n1 + 1] = wv.y * X[16 * (n1 + 1) + (n1 + 1 >= 10) + (n1 + 1 >= 20)];
            }
            codelets::w_pass1<float>(y, re, im);
            sm.ex0[role][lane] = re[0];
#pragma unroll
            for (int k1 = 1; k1 < 13; ++k1) {
                // the codelet leaves im[k1] short of a constant factor (its last butterfly's sine).  Folding the factor into
                // a four-entry twiddle costs more in constant loads than the multiply does (measured: 130 vs 126 us on cfg2)
                const float2 t = cw_tw[role][k1];
                const float imk = im[k1] * float(codelets::w_pass1_im_scale_of(k1));
                sm.ex[k1 - 1][role][lane] = make_float2(fmaf(re[k1], t.x, -(imk * t.y)), fmaf(re[k1], t.y, imk * t.x));
            }
        };
        pass1(warp);
        pass1(warp + kWarps);
        __syncthreads();                            // exchange complete; xs is dead, its storage becomes the power spectrum

        // ---- pass 2 (row = k1 = warp, warp + 8; rows 0..12) + power ----
#pragma unroll 1
        for (int row = warp; row < 13; row += kWarps) {
            if (row == 0) {
                float a[16], er[9], ei[9];
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) a[n2] = sm.ex0[n2][lane];
                codelets::w_pass2_edge<float>(a, er, ei);
#pragma unroll
                for (int k2 = 0; k2 < 9; ++k2) sm.u.P[25 * k2][lane] = fmaf(er[k2], er[k2], ei[k2] * ei[k2]);
            } else {
                float xr[16], xi[16], yr[16], yi[16];
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) {
                    const float2 v = sm.ex[row - 1][n2][lane];
                    xr[n2] = v.x; xi[n2] = v.y;
                }
                codelets::dft16<float>(xr, xi, yr, yi);
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2)
                    sm.u.P[k2 < 8 ? row + 25 * k2 : 400 - row - 25 * k2][lane] = fmaf(yr[k2], yr[k2], yi[k2] * yi[k2]);
            }
        }
        // the next tile (of this item or the next): descriptor, then its PCM (staging has been free since the layout pass).
        // 13 rows over 8 warps: warps 5..7 have one row where the others have two, so the last warp has time for it
        if (tid == kThreads - 32) {
            next_tile(sm.desc[slot], sm.desc[slot ^ 1]);
            prefetch(sm.desc[slot ^ 1]);
        }
        __syncthreads();

        // ---- mel + log10: lane <-> frame, so the stores along t coalesce ----
        {
            const float* Pl = &sm.u.P[0][lane];
            const int t = t0 + lane;
#if STX_W_MEL_EXACT
            // mel bin warp + 8 h + 16 i = bin `warp` of half slot 2 i + h
            {
                float v[5];
                v[0] = mel_slot10<0>(Pl, sm.melw, sm.melfirst, warp);
                v[1] = mel_slot10<2>(Pl, sm.melw, sm.melfirst, warp);
                v[2] = mel_slot10<4>(Pl, sm.melw, sm.melfirst, warp);
                v[3] = mel_slot10<6>(Pl, sm.melw, sm.melfirst, warp);
                v[4] = mel_slot10<8>(Pl, sm.melw, sm.melfirst, warp);
                float* const po = out_b + (size_t)warp * T + t;      // mel bin `warp`, frame t; bin m is m * T floats further
#if STX_W_LAZY_CLAMP
                float lo5 = fminf(fminf(fminf(v[0], v[1]), fminf(v[2], v[3])), v[4]);
                if (t < t_end) {
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        po[(size_t)(16 * i) * T] = fmaf(v[i], 0.25f, 1.0f);      // = (x + 4) / 4 in float32, bit for bit
                        run_max = fmaxf(run_max, v[i]);
                    }
                }
#else
                if (t < t_end) {
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        po[(size_t)(16 * i) * T] = v[i];
                        run_max = fmaxf(run_max, v[i]);
                    }
                }
#endif
                v[0] = mel_slot10<1>(Pl, sm.melw, sm.melfirst, warp);
                v[1] = mel_slot10<3>(Pl, sm.melw, sm.melfirst, warp);
                v[2] = mel_slot10<5>(Pl, sm.melw, sm.melfirst, warp);
                v[3] = mel_slot10<7>(Pl, sm.melw, sm.melfirst, warp);
                v[4] = mel_slot10<9>(Pl, sm.melw, sm.melfirst, warp);
#if STX_W_LAZY_CLAMP
                lo5 = fminf(lo5, fminf(fminf(fminf(v[0], v[1]), fminf(v[2], v[3])), v[4]));
                if (t < t_end) {
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        po[(size_t)(8 + 16 * i) * T] = fmaf(v[i], 0.25f, 1.0f);
                        run_max = fmaxf(run_max, v[i]);
                    }
                    cur.fmin_b[(size_t)warp * T + t] = lo5;
                }
#else
                if (t < t_end) {
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        po[(size_t)(8 + 16 * i) * T] = v[i];
                        run_max = fmaxf(run_max, v[i]);
                    }
                }
#endif
            }
#else
#pragma unroll 1
            for (int role = warp; role < 16; role += kWarps) {
                float v[5];
                v[0] = mel_slot<0>(Pl, sm.melw, sm.melfirst, role);
                v[1] = mel_slot<1>(Pl, sm.melw, sm.melfirst, role);
                v[2] = mel_slot<2>(Pl, sm.melw, sm.melfirst, role);
                v[3] = mel_slot<3>(Pl, sm.melw, sm.melfirst, role);
                v[4] = mel_slot<4>(Pl, sm.melw, sm.melfirst, role);
                if (t < t_end) {
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        out_b[(size_t)(role + 16 * i) * T + t] = v[i];
                        run_max = fmaxf(run_max, v[i]);
                    }
                }
            }
#endif
        }
        __syncthreads();                            // the power spectrum is consumed: the next layout pass may overwrite it
        }
        if (cur.last) {                             // the item is complete: one atomic per warp publishes the clip's running maximum
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) run_max = fmaxf(run_max, __shfl_xor_sync(0xffffffffu, run_max, o));
            if (lane == 0) atomicMax(cur.cmax, max_key(run_max));
            run_max = __int_as_float(0xff800000);
        }
        slot ^= 1;
    }
}

// max(x, max - 8), (x + 4) / 4 in place; mask[b, j] = (160 j < len)
__global__ void __launch_bounds__(256)
w_finish(const int* __restrict__ lengths, const unsigned* __restrict__ clip_max, int n_samples,
         float* __restrict__ out, int* __restrict__ mask, const float* __restrict__ frame_min) {
    const int b = blockIdx.y;
    const int T = n_samples / kHop;
    const float lo = max_unkey(clip_max[b]) - 8.0f;
    const size_t total = (size_t)kMel * T;
    float* o = out + (size_t)b * total;
#if STX_W_LAZY_CLAMP
    // one thread per frame: the frame's smallest log10 (over the eight warps' minima) against the clip's floor; a frame below it
    // is clamped in place in the scaled domain.  (x + 4) / 4 is monotone, so max((x + 4) / 4, (lo + 4) / 4) has the bits of
    // (max(x, lo) + 4) / 4
    {
        const float* fm = frame_min + (size_t)b * kWarps * T;
        const float ylo = fmaf(lo, 0.25f, 1.0f);
        const float kNegInf = __int_as_float(0xff800000);
        // a frame of a padding-only tile (marker -inf in warp 0's minimum) was not written at all: every bin is log10 of the
        // 1e-10 floor there, so its final value is the same for all 80 bins
        const float ypad = fmaf(fmaxf(log10_pos(1e-10f), lo), 0.25f, 1.0f);
        if ((T & 3) == 0 && ((reinterpret_cast<unsigned long long>(fm) | reinterpret_cast<unsigned long long>(o)) & 15ull) == 0) {
            // four frames per thread: eight 16-byte loads
            const float4* fm4 = reinterpret_cast<const float4*>(fm);
            const int T4 = T >> 2;
            for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < T4; q += gridDim.x * blockDim.x) {
                float4 m = fm4[q];
                const bool px = m.x == kNegInf, py = m.y == kNegInf, pz = m.z == kNegInf, pw = m.w == kNegInf;
#pragma unroll
                for (int w = 1; w < kWarps; ++w) {
                    const float4 v = fm4[(size_t)w * T4 + q];
                    m.x = fminf(m.x, v.x); m.y = fminf(m.y, v.y); m.z = fminf(m.z, v.z); m.w = fminf(m.w, v.w);
                }
                float4* o4 = reinterpret_cast<float4*>(o) + q;
                if (px && py && pz && pw) {
                    const float4 v = make_float4(ypad, ypad, ypad, ypad);
#pragma unroll 8
                    for (int r = 0; r < kMel; ++r) o4[(size_t)r * T4] = v;
                } else if (fminf(fminf(m.x, m.y), fminf(m.z, m.w)) < lo) {
                    // (a frame whose minimum is not below the floor has every value at or above ylo: the max leaves it as it is,
                    // so the four frames go through it together -- 16-byte accesses, consecutive threads on consecutive quads)
#pragma unroll 8
                    for (int r = 0; r < kMel; ++r) {
                        float4 v = o4[(size_t)r * T4];
                        v.x = px ? ypad : fmaxf(v.x, ylo); v.y = py ? ypad : fmaxf(v.y, ylo);
                        v.z = pz ? ypad : fmaxf(v.z, ylo); v.w = pw ? ypad : fmaxf(v.w, ylo);
                        o4[(size_t)r * T4] = v;
                    }
                }
            }
        } else {
            for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
                float m = fm[t];
                const bool pad = m == kNegInf;
#pragma unroll
                for (int w = 1; w < kWarps; ++w) m = fminf(m, fm[(size_t)w * T + t]);
                if (pad) {
#pragma unroll 8
                    for (int r = 0; r < kMel; ++r) o[(size_t)r * T + t] = ypad;
                } else if (m < lo) {
#pragma unroll 8
                    for (int r = 0; r < kMel; ++r) o[(size_t)r * T + t] = fmaxf(o[(size_t)r * T + t], ylo);
                }
            }
        }
    }
#else
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
#endif
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
        static float win[16][26] = {};
        static float2 tw[16][16];
        for (int n2 = 0; n2 < 16; ++n2)
            for (int n1 = 0; n1 < 25; ++n1) win[n2][n1] = (float)w[16 * n1 + n2];
        for (int n2 = 0; n2 < 16; ++n2)
            for (int k1 = 0; k1 < 16; ++k1) {
                const double ang = -2.0 * M_PI * double((n2 * k1) % 400) / 400.0;
                tw[n2][k1] = make_float2((float)std::cos(ang), (float)std::sin(ang));
            }
        STX_CUDA(cudaMemcpyToSymbol(cw_win, win, sizeof(win)));
        STX_CUDA(cudaMemcpyToSymbol(cw_tw, tw, sizeof(tw)));
        const std::vector<double>& fb = w_mel();
#if STX_W_MEL_EXACT
        for (int i = 0; i < kMelWeights; ++i) h.melw[i] = 0.0f;
        for (int m = 0; m < kMel; ++m) {
            const int j = m / 8, wrp = m % 8, L = mel_len10(j), PAD = mel_pad10(j);
            int lo = -1, hi = -1;
            for (int k = 0; k < kBins; ++k)
                if (fb[size_t(k) * kMel + m] != 0.0) { if (lo < 0) lo = k; hi = k; }
            if (lo < 0 || hi - lo + 1 > L) { set_error("mel filter %d does not fit its half slot", m); return STX_EINVAL; }
            if (lo + L > kBins) lo = kBins - L;
            h.melfirst[m] = lo;
            for (int q = 0; q < L; ++q) h.melw[mel_off10(j) + wrp * PAD + q] = float(fb[size_t(lo + q) * kMel + m]);
        }
#else
        for (int m = 0; m < kMel; ++m) {
            const int slot = m / 16, wrp = m % 16, L = mel_len(slot);
            int lo = -1, hi = -1;
            for (int k = 0; k < kBins; ++k)
                if (fb[size_t(k) * kMel + m] != 0.0) { if (lo < 0) lo = k; hi = k; }
            if (lo < 0 || hi - lo + 1 > L) { set_error("mel filter %d does not fit its slot", m); return STX_EINVAL; }
            if (lo + L > kBins) lo = kBins - L;
            h.melfirst[m] = lo;
            for (int q = 0; q < L; ++q) h.melw[mel_off(slot) + wrp * L + q] = float(fb[size_t(lo + q) * kMel + m]);
        }
#endif
        WTables* d = nullptr;
        STX_CUDA(cudaMalloc(&d, sizeof(WTables)));
        STX_CUDA(cudaMemcpy(d, &h, sizeof(WTables), cudaMemcpyHostToDevice));
        STX_CUDA(cudaFuncSetAttribute(w_frames<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        STX_CUDA(cudaFuncSetAttribute(w_frames<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        STX_CUDA(cudaFuncSetAttribute(w_frames<kStockSamples>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        STX_CUDA(cudaFuncSetAttribute(w_frames<kStockSamples>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        g_tab[dev] = d;
    }
    *out = g_tab[dev];
    return 0;
}

// Frames per CTA (a multiple of the 32-frame tile), chosen like fbank_k.cu's pick_chunk: minimise waves x tiles.
inline int pick_chunk(int B, int frames, int sms) {
    int best = 64;
    long long best_cost = -1;
    for (int chunk = 64; chunk <= 512; chunk += kTile) {
        const long long ctas = (long long)B * ((frames + chunk - 1) / chunk);
        const long long cost = ((ctas + 2 * sms - 1) / (2 * sms)) * (chunk / kTile);      // two CTAs per SM
        if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = chunk; }
    }
    return best;
}

}  // namespace
}  // namespace stx

extern "C" {

int stx_logmel_w_workspace(int B, int n_samples, size_t* bytes) {
    if (B < 0 || n_samples < 0 || !bytes) { stx::set_error("stx_logmel_w_workspace: bad argument"); return STX_EINVAL; }
    // the per-clip maxima, then (STX_W_LAZY_CLAMP) the per-frame, per-warp minima [B][8][n_samples / 160]
    *bytes = ((size_t(B) * sizeof(float) + 255) & ~size_t(255)) + 256 +
             ((size_t(B) * stx::kWarps * (size_t)(n_samples / stx::kHop) * sizeof(float) + 255) & ~size_t(255));
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
    unsigned* clip_max = static_cast<unsigned*>(d_ws);
    float* frame_min = reinterpret_cast<float*>(static_cast<unsigned char*>(d_ws) + ((size_t(B) * sizeof(float) + 255) & ~size_t(255)) + 256);
    const int T = n_samples / kHop;
    STX_CUDA(cudaMemsetAsync(clip_max, 0, size_t(B) * sizeof(unsigned), st));
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        STX_CUDA(cudaGetDevice(&dev));
        STX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int chunk_frames = pick_chunk(B, T, sms);
    if (const char* e = std::getenv("STX_W_CHUNK")) {        // development: force the frames per CTA (a multiple of 32, >= 32)
        const int v = std::atoi(e);
        if (v >= kTile && v % kTile == 0) chunk_frames = v;
    }
    const int chunks = (T + chunk_frames - 1) / chunk_frames;
    const int grid = (int)std::min<long long>(2LL * sms, (long long)B * chunks);
    if (n_samples == kStockSamples)
        STX_LAUNCH(w_frames<kStockSamples>, dim3(grid), dim3(kThreads), sizeof(Smem), st,
                   d_pcm, reinterpret_cast<const long long*>(d_offsets), d_lengths, d_peak, tab, B, n_samples, chunk_frames,
                   chunks, d_out, clip_max, frame_min);
    else
        STX_LAUNCH(w_frames<0>, dim3(grid), dim3(kThreads), sizeof(Smem), st,
                   d_pcm, reinterpret_cast<const long long*>(d_offsets), d_lengths, d_peak, tab, B, n_samples, chunk_frames,
                   chunks, d_out, clip_max, frame_min);
#if STX_W_LAZY_CLAMP
    const int gx = std::max(1, ((T + 3) / 4 + 255) / 256);  // one thread per four frames
#else
    const size_t total = (size_t)kMel * T;
    const int gx = (int)std::max<size_t>(1, std::min<size_t>((total / 4 + 255) / 256, 64));
#endif
    STX_LAUNCH(w_finish, dim3(gx, B), dim3(256), 0, st, d_lengths, clip_max, n_samples, d_out, d_mask, (const float*)frame_min);
    return 0;
}

}  // extern "C"
