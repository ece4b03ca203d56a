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
// Kernels:
//   k_frames_duo  (shipped) persistent, one CTA per SM, two independent 8-warp groups, each streaming through its own list of
//                 (clip, chunk of frames) work items: PCM -> raw log-mel (written in place into the output tensor, whose
//                 [T_pad/2,160] rows are exactly [T_pad,80] rows) + per-item per-bin (sum, sum of squares) partials in fixed
//                 point (int64).  Same phases as k_frames below; the exchange goes through a 64 KB buffer in two halves with
//                 the second half stashed in tensor memory (see the comment above the kernel)
//   k_frames      its single-group predecessor, one CTA per (clip, chunk): kept for A/B runs (STX_K_SINGLE=1) and as the
//                 readable statement of the phases
//   k_schedule    ragged batches only: compacts the non-empty work items on the device (k_frames_duo is its programmatic
//                 dependent); skipped when the caller promises a uniform batch (negated max_length)
//   k_normalize   every CTA sums its clip's partials (integers: same result in every CTA) -> mean, 1/sqrt(var_ddof1 + 1e-7);
//                 then in-place CMVN, padding rows, attention mask (the extractor's int32 mask, or the trainer collate's int64
//                 mask with zero rows past the clip: R/training/trainer_unfreeze.py:898-908)
//   k_norm_ln_split  k_normalize fused with the encoder's LayerNorm + TF32 split (stx_fbank_k_projection)
//
// k_frames keeps ONE FRAME PER LANE: a tile is 32 consecutive frames of a clip and the 16 warps of the
// CTA split each frame's 512-point real FFT (n = 16 n1 + n2, k = k1 + 32 k2) between them, so every
// per-warp quantity that is not data (window, twiddles, mel weights) is warp-uniform and comes from the
// constant bank or a shared-memory broadcast, and every shared-memory access is lane-contiguous
// (conflict-free by construction):
//
//   stage    one cp.async.bulk (TMA 1-D bulk copy) per tile lands the tile's 5364 raw samples in shared memory
//            on an mbarrier while the previous tile is transformed (clips that are not 16-byte aligned, and the
//            <= 3 tail samples of a clip, are read with plain loads)
//   convert  every sample is converted to float64 ONCE: d[i] = x[i] - 0.97 x[i-1] (the pre-emphasis of every
//            frame at once; the frame mean enters after the FFT), stored in rows of 161 doubles (odd stride:
//            lane f reads row f + const without bank conflicts)
//   pass 1   warp n2: y[n1] = W[i] d[i], i = 16 n1 + n2 (W = 2^15 * Povey), real DFT-32 over n1 in
//            registers (codelets.cuh, generated), times W512^(n2 k1) -> shared [k1][n2][lane]
//   pass 2   warp k1: complex DFT-16 over n2 -> X[k1 + 32 k2] (k2 >= 8 are the mirrored bins 512 - k, same power);
//            warp 0 does the two real rows k1 = 0 and k1 = 16
//   DC       the reference subtracts the frame mean m before pre-emphasis; with c = 0.03 m that is
//            y - c W, i.e. X - c * FFT(W) by linearity: 2 FMAs per bin with a constant table;
//            400 c = sum(d) - 0.97 (x[399] - x[-1]) from a two-level sum of the per-warp sums of d
//   power    |X|^2 in float64, rounded to float32 -> shared [bin][lane]
//   mel      warp w: mel bins w, w + 16, ...: filters padded to a fixed length per slot (no loop, no
//            metadata decode) -> ln -> staged rows -> coalesced stores + float64 sum / sum of squares per bin
#include "stx_common.h"
#include "codelets.cuh"
#include "mel_k.cuh"
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <type_traits>

// Build-time switches for A/B runs (tools/ab_build.py); the defaults are the shipped configuration.
#ifndef STX_K_MEL_GEN
#define STX_K_MEL_GEN 0          // mel stage of k_frames_duo (cfg2, us per launch of k_frames_duo, profiles/r02_k_variants.md):
                                 // 0: filter-walking, weights and first bins in shared memory (round 1)            217.3
                                 // 1: generated bin-walking stage (mel_k.cuh): a third of the shared-memory reads and 8 % fewer
                                 //    instructions, but eight divergent code paths per group -- instruction-fetch stalls go
                                 //    from 0.13 to 1.01 per issue                                                  236.8
                                 // 2: filter-walking with the (warp-uniform) weights in the constant bank: 430 fewer
                                 //    shared-memory wavefronts per tile, 54 more uniform loads                      220.7
#endif
#ifndef STX_K_CLIPSTATS
#define STX_K_CLIPSTATS 0        // 1: the group that completes a clip's last work item (atomic counter per clip) reduces the clip's
                                 //    statistics inside k_frames_duo, so that k_normalize starts streaming at once: k_normalize
                                 //    23.0 -> 20.8 us, but k_frames_duo 220.7 -> 236.2 us (a fence, two group barriers and an L2
                                 //    atomic per work item sit on the group's critical path).  0: every CTA of k_normalize sums the
                                 //    clip's partials itself (round 1)
#endif
#ifndef STX_K_CVT_SPLIT
#define STX_K_CVT_SPLIT 0        // 1: the conversion pass of k_frames_duo widens one of its two float32 operands per sample on the
                                 //    integer pipe (exponent re-bias + shifts) and the other with F2F, instead of both on the
                                 //    quarter-rate XU pipe
#endif
#ifndef STX_K_BAR2_LATE
#define STX_K_BAR2_LATE 0        // 1: the "H2 complete" barrier of k_frames_duo moves behind the first half of pass 2, so that the
                                 //    shared-memory stores of the second exchange half drain under that half's FP64 work
#endif
#ifndef STX_K_NORM_PDL
#define STX_K_NORM_PDL 0         // 1: k_normalize is launched as the programmatic dependent of k_frames_duo (no measurable gain:
                                 //    the frames kernel owns every SM until its last CTA retires)
#endif

#ifndef STX_K_ILP
#define STX_K_ILP 3              // bit mask over the phases of k_frames_duo that are bound by latency, not by a pipe (cfg2, us per launch):
                                 //   1  mel: all ten filters of a warp are accumulated in registers before the first result is stored
                                 //      (a shared-memory store between two filters makes ptxas serialise their load -> FMA chains),
                                 //      lg2.approx.ftz (drops the denormal fix-up code)                         218.8 -> 216.2
                                 //   2  store + statistics: full tiles take a branch-free path, so the 11 load -> convert -> split ->
                                 //      add chains of a thread overlap instead of running as 11 guarded blocks   218.8 -> 212.6
                                 //   4  conversion: runs of 25 consecutive samples per thread (26 float32 -> float64 conversions for
                                 //      25 outputs instead of 50).  SLOWER (226.4): not adopted
                                 //   3 = 1 + 2 (shipped)                                                         218.8 -> 209.6
#endif
#ifndef STX_K_POW_PRESCALE
#define STX_K_POW_PRESCALE 1     // 1: the window table carries an extra factor 2^-448 (exact), so that the power spectrum comes out
                                 //    scaled by 2^-896 and its float64 exponent field IS the float32 exponent field of the unscaled
                                 //    value: the conversion on the integer pipe is one funnel shift instead of clamp + re-bias +
                                 //    shift (64 instructions per warp and tile).  (Float64 transform only: forced off in the
                                 //    float32 builds of the precision study.)
#endif
#ifndef STX_K_STASH_PAIRS
#define STX_K_STASH_PAIRS 1      // 1: the tensor-memory stash is written one double per tcgen05.st (.x2) instead of one .x32
#endif
#ifndef STX_K_NORM_COLS
#define STX_K_NORM_COLS 1        // 1: k_normalize's threads own a column quad each (statistics in registers, no division); 0: round 1
#endif
#ifndef STX_K_CVT_LINEAR
#define STX_K_CVT_LINEAR 1       // 1: the conversion pass addresses its 23 half rows as base + constant (see there); 0: round-1 form
#endif
#ifndef STX_K_CVT_ROWS
#define STX_K_CVT_ROWS 0         // 1: the conversion pass of k_frames_duo walks whole rows with 160 threads (constant strides)
#endif
#ifndef STX_K_POWER_F32
#define STX_K_POWER_F32 0        // 1: the DC-corrected spectrum value is rounded to float32 (two F2F) and squared on the FP32 pipe (like
                                 //    the reference, which rounds the spectrum to complex64 before |.|^2) instead of |.|^2 in float64
                                 //    converted on the integer pipe: 2 FP64 + 2 XU + 2 FP32 instead of 4 FP64 + 3 ALU instructions per
                                 //    bin.  Slower (222.3 vs 218.1 us): an F2F occupies the XU pipe of its scheduler for 8-16 cycles
                                 //    (tools/microbench_coissue.cu), and 94 more of them per warp and tile cost more than the 64 FP64
                                 //    instructions they replace
#endif
#ifndef STX_K_MEL_EXACT
#define STX_K_MEL_EXACT 1        // 1 (shipped): the mel stage of k_frames_duo walks ten half slots of 8 filters (one per warp of a
                                 //    group) whose lengths are the longest filter of the half slot (2, 3, 3, 4, 5, 6, 8, 10, 13, 16
                                 //    bins: 560 bin reads and FMAs per frame) instead of five slots of 16 padded to 4, 4, 8, 12, 16
                                 //    (704): 206.4 -> 201.4 us on cfg2, bit-identical (the dropped products had zero weights)
#endif
#ifndef STX_K_H2_TMEM
#define STX_K_H2_TMEM 0          // 1: the rows of the second exchange half that a warp needs from the two warps sharing its tensor-memory
                                 //    lane quarter (4 of the 16 roles, itself included) are read straight from their stash
                                 //    (tcgen05.ld) instead of going through shared memory: a quarter of the H2 stores and loads
                                 //    (256 of ~5500 shared-memory wavefronts per tile).  Bit-identical and no faster (201.2 against
                                 //    200.5 us on cfg2): the exchange phases are not bound by their wavefront count
#endif
#ifndef STX_K_LN_FAST
#define STX_K_LN_FAST 1          // 1 (shipped): ln of a mel energy as lg2.approx of the whole value times ln 2 (2 instructions instead
                                 //    of 8): 201.1 -> 197.7 us on cfg2.  The log2 is a float32 of magnitude up to 64, so the raw log-mel
                                 //    is good to 2 float32 ulps (3.8e-6; 7.6e-6 for full-scale int16-valued input) instead of 1; the
                                 //    normalised features of cfg2 move from 1.31e-5 to 1.59e-5 of the 1e-4 bar (all 64 clips)
#endif
#ifndef STX_K_STATS_TILE
#define STX_K_STATS_TILE 1       // 1 (shipped, 197.8 -> 196.4 us on cfg2): a thread sums its <= 11 values of a tile (and their squares) in float64 first and converts the two
                                 //    sums to fixed point once per tile, instead of converting every value (5 FP64 + 6 integer
                                 //    instructions per value become 2 FP64).  The tile grid of a clip does not depend on the batch, so
                                 //    the statistics stay batch-invariant; they differ from the per-value form in the last bits
#endif
#ifndef STX_K_SOLO
#define STX_K_SOLO 0             // 1 (experiment): only group 0 of every CTA works
#endif
#ifndef STX_K_TRACE
#define STX_K_TRACE 0            // 1: thread 0 of every group of k_frames_duo adds the cycles between its group barriers to per-phase
                                 //    counters (stx_debug_ktrace reads and clears them; tools/k_phase_trace.py)
#endif

namespace stx {
namespace {

#if STX_K_TRACE
__device__ unsigned long long g_ktrace[64];
__device__ __forceinline__ unsigned clock_lo() { unsigned c; asm volatile("mov.u32 %0, %%clock;" : "=r"(c)); return c; }
#endif
#if STX_K_TRACE == 1
#define KTRACE(i) do { if (tl == 0) { const unsigned now__ = clock_lo(); sg.tacc[i] += now__ - sg.tlast; sg.tlast = now__; } } while (0)
#define KTRACE_PRE(i) do { } while (0)
#elif STX_K_TRACE == 2
// per-warp BUSY cycles of each phase (from the warp's release by the previous barrier to its arrival at the next one)
#define KTRACE_PRE(i) do { if (lane == 0) sg.tw[i][w8] += clock_lo() - sg.tl[w8]; } while (0)
#define KTRACE(i) do { if (lane == 0) sg.tl[w8] = clock_lo(); } while (0)
#else
#define KTRACE(i) do { } while (0)
#define KTRACE_PRE(i) do { } while (0)
#endif

constexpr int kFrame = STX_K_FRAME;
constexpr int kHop = STX_K_HOP;
constexpr int kMel = STX_K_NMEL;
constexpr int kTile = 32;                        // frames per tile = lanes
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;            // 512
constexpr int kTileSamples = (kTile - 1) * kHop + kFrame;   // 5360
constexpr int kLead = 4;                         // the staged range starts 4 samples early (x[-1], 16-byte alignment)
constexpr int kStage = kLead + kTileSamples + 4; // 5368 floats: stage[q] = x[160 t0 - 4 + q]
constexpr int kDRow = kHop + 1;                  // d rows: 160 doubles + 1 pad
constexpr int kDBuf = 34 * kDRow;                // 33.5 hops per tile
constexpr int kMinChunk = 64;                    // smallest chunk of frames per CTA (sizes the partials workspace)
constexpr float kMelFloor = 1.192092955078125e-07f;
constexpr int kOutRow = kMel + 1;                // staged log-mel rows: odd stride, lane <-> row is conflict-free
constexpr int kStatThreads = 6 * kMel;           // 480 threads store 6 rows of 80 bins per step
// Statistics are accumulated in FIXED POINT (int64), so they do not depend on how a clip is cut into chunks,
// tiles or row groups: a clip's features are bit-identical whatever batch it is in.  Adding 1.5 * 2^(52 - s) to
// a double rounds it to a multiple of 2^-s and leaves that integer in the low mantissa bits.  x is a float32
// with |x| < 2^7 (a natural log), so x is exact on the 2^-32 grid down to |x| = 2^-9 and x^2 (exact in
// float64) is split into a multiple of 2^-20 plus a remainder below 2^-21 kept on the 2^-56 grid; 2^24 frames
// per clip fit in 63 bits.  (A single 2^-28 grid for x^2 is not enough: a 2-frame clip has variances ~1e-7.)
constexpr double kFix1 = 1.5 * 1048576.0;        // 1.5 * 2^20: grid 2^-32 (sum x)
constexpr double kFixH = 1.5 * 4294967296.0;     // 1.5 * 2^32: grid 2^-20 (sum x^2, high part)
constexpr double kFixL = 1.5 * 0.0625;           // 1.5 * 2^-4: grid 2^-56 (sum x^2, low part)
constexpr int kStatWords = 3 * kMel;             // per chunk: sum x | sum x^2 high | sum x^2 low
// mel slot i holds bins 16 i .. 16 i + 15 (one per warp), every filter of a slot padded to the same length
__host__ __device__ constexpr int mel_len(int slot) { return slot == 0 ? 4 : slot == 1 ? 4 : slot == 2 ? 8 : slot == 3 ? 12 : 16; }
__host__ __device__ constexpr int mel_off(int slot) { return slot == 0 ? 0 : mel_off(slot - 1) + 16 * mel_len(slot - 1); }
constexpr int kMelWeights = mel_off(5);          // 704
// (STX_K_MEL_EXACT) half slot j holds mel bins 8 j .. 8 j + 7 (one per warp of a group): mel_len10 bins are read, the weights are
// stored padded to a multiple of four (aligned float4 loads)
__host__ __device__ constexpr int mel_len10(int j) {
    return j == 0 ? 2 : j == 1 ? 3 : j == 2 ? 3 : j == 3 ? 4 : j == 4 ? 5 : j == 5 ? 6 : j == 6 ? 8 : j == 7 ? 10 : j == 8 ? 13 : 16;
}
__host__ __device__ constexpr int mel_pad10(int j) { return (mel_len10(j) + 3) & ~3; }
__host__ __device__ constexpr int mel_off10(int j) { return j == 0 ? 0 : mel_off10(j - 1) + 8 * mel_pad10(j - 1); }
constexpr int kMelWeights10 = mel_off10(10);     // 672
static_assert(kMelWeights10 <= kMelWeights, "the half-slot table fits the shared-memory array of the slot table");

// warp-uniform tables (constant bank)
__constant__ double  c_win[16][25];              // [n2][n1] = W[16 n1 + n2], W = 2^15 * Povey
__constant__ double2 c_tw[16][16];               // [n2][k1] = W512^(n2 k1)
__constant__ double2 c_wh[16][16];               // [k1][k2] = FFT512(W)[k1 + 32 k2]; row 0: [k2] = bin 32 k2 (k2 = 1..7), [8 + k2] = bin 16 + 32 k2

struct KTables {
    float melw[kMelWeights];     // [slot][warp][mel_len(slot)]
    int   melfirst[kMel];        // first FFT bin of the padded filter of mel bin m
    float melw10[kMelWeights];   // (STX_K_MEL_EXACT) [half slot][warp of a group][mel_pad10(half slot)]
    int   melfirst10[kMel];
};

// Precision study (north_star: "the choice evidenced by ncu"; profiles/r02_k_precision.md).  The single-group kernel k_frames
// (STX_K_SINGLE=1) can be built with its first pass (pre-emphasis, window, real DFT-32, twiddles) and / or its second pass
// (exchange, DFT-16, DC correction, power) in float32: tools/ab_build.py f32=-DSTX_K_P1_F32=1,-DSTX_K_P2_F32=1 etc.
// The shipped kernels are float64 in both (see the header comment for why).
#ifndef STX_K_P1_F32
#define STX_K_P1_F32 0
#endif
#ifndef STX_K_P2_F32
#define STX_K_P2_F32 0
#endif
template <bool F32> struct real_of { using type = double; using pair = double2; };
template <> struct real_of<true> { using type = float; using pair = float2; };
using r1_t = real_of<STX_K_P1_F32 != 0>::type;    // pass-1 arithmetic
using r2_t = real_of<STX_K_P2_F32 != 0>::type;    // exchange and pass-2 arithmetic
using r2x2_t = real_of<STX_K_P2_F32 != 0>::pair;
#if STX_K_P1_F32 || STX_K_P2_F32
__constant__ float  c_win_f[16][25];
__constant__ float2 c_tw_f[16][16];
__constant__ float2 c_wh_f[16][16];
#endif

struct Smem {
    r2x2_t  ex[15][16][kTile];   // pass-1 output rows k1 = 1..15: [k1 - 1][n2][lane]; later aliased by the staged
                                 // log-mel rows and the statistics reduction
    r2_t    ex0[16][kTile];      // row k1 = 0  (real)
    r2_t    ex16[16][kTile];     // row k1 = 16 (real)
    union {
        r1_t  d[kDBuf];          // pre-emphasised samples of the tile (float64), rows of 161
        float P[256][kTile];     // power spectrum [bin][lane] (pass 2 onwards); row 0 is zeroed every tile
    } u;
    float   stage[kStage];       // raw PCM of the next tile, landed by cp.async.bulk
    r1_t    psum[kTile][17];     // per-warp sums of d, [lane][n2]
    r2_t    cval[kTile];         // 0.03 * mean of each frame
    r1_t    xb[kTile];           // 0.97 (x[399] - x[-1]) of each frame
    float   melw[kMelWeights];
    int     melfirst[kMel];
    unsigned long long mbar;     // completion barrier of the bulk copy
    unsigned long long cbar;     // "all 32 values of cval are written" (one arrival per warp)
};
static_assert(sizeof(Smem) <= 227 * 1024, "one CTA per SM must fit in 227 KB");
static_assert(offsetof(Smem, stage) % 16 == 0, "bulk-copy destination alignment");
static_assert(kTile * kOutRow * sizeof(float) <= sizeof(r2x2_t) * 15 * 16 * kTile, "staged rows alias the exchange area");
static_assert(3 * kStatThreads * sizeof(unsigned long long) <= sizeof(r2x2_t) * 15 * 16 * kTile, "statistics reduction aliases the exchange area");

// ln(x) for normal positive x (the mel floor guarantees it): exponent + MUFU.LG2 of the mantissa.
// |error| <= ~0.6 ulp of the result for results of magnitude 10..30 (the mantissa's log2 is in [0, 1), where
// lg2.approx is accurate to 2^-22 absolute), i.e. as good as logf at a third of the instructions.
__device__ __forceinline__ float ln_pos(float x) {
#if STX_K_LN_FAST
    // (A/B) log2 of the whole value in one MUFU: the result is rounded to float32 at magnitude up to 64, i.e. up to 3.8e-6 in
    // log2 (2.6e-6 in ln) where the split form keeps ~1e-7
    float l2x;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2x) : "f"(x));
    return l2x * 0.69314718055994530942f;
#endif
    const int bits = __float_as_int(x);
    const float e = (float)((bits >> 23) - 127);
    const float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
    float l2;
#if (STX_K_ILP & 1)
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(m));     // m is in [1, 2): .ftz only removes the denormal-input fix-up code
#else
    asm("lg2.approx.f32 %0, %1;" : "=f"(l2) : "f"(m));
#endif
    constexpr float ln2_hi = 0.693145751953125f;          // 16 significant bits: e * ln2_hi is exact
    constexpr float ln2_lo = 1.42860682030941723212e-6f;
    constexpr float ln2 = 0.69314718055994530942f;
    return fmaf(e, ln2_hi, fmaf(l2, ln2, e * ln2_lo));
}

// float64 -> float32 of a power-spectrum value on the integer pipe.  F2F.F32.F64 is a quarter-rate XU instruction and the
// frame-per-lane kernel runs its phases in lockstep, so the 256 conversions per frame cost ~5 % of the tile time.  The value
// is a non-negative finite double far below 2^128: re-bias the exponent (1023 -> 127) and funnel-shift the top 23 mantissa
// bits into place.  Truncation instead of round-to-nearest (<= 1 float32 ulp, 1.2e-7 relative on a value whose LOGARITHM
// is compared at 1e-4); anything below 2^-126 (exact zeros: digital silence) becomes a denormal <= 7 * 2^-149, i.e. zero for
// the mel floor that follows.
__device__ __forceinline__ float power_to_f32(float p) { return p; }
__device__ __forceinline__ float power_to_f32(double p) {
    const unsigned hi = (unsigned)__double2hiint(p), lo = (unsigned)__double2loint(p);
#if STX_K_POW_PRESCALE && !STX_K_P1_F32 && !STX_K_P2_F32
    // p is the power scaled by 2^-896 (see kWinScale): exponent field = float32's, anything below 2^-126 is a float64 denormal
    // or zero and becomes a float32 denormal <= 7 * 2^-149 or zero by itself
    return __uint_as_float(__funnelshift_l(lo, hi, 3));
#else
    const unsigned h = max(hi, 0x38000000u) - 0x38000000u;
    return __uint_as_float(__funnelshift_l(lo, h, 3));
#endif
}

// float32 -> float64 on the integer pipe (exact for normal numbers: re-bias the exponent by 896, shift the mantissa).  Zeros
// and denormals come out as values below 2^-126 instead of exactly: at most 1.2e-38 in a sample that is then scaled by <= 2^15
// and squared, i.e. far below the mel floor -- the output bits are the same as with an exact conversion.
__device__ __forceinline__ double f32_to_f64_int(float x) {
    const unsigned b = __float_as_uint(x);
    const unsigned hi = (((b >> 3) & 0x0fffffffu) + 0x38000000u) | (b & 0x80000000u);
    return __hiloint2double((int)hi, (int)(b << 29));
}

// ---- mbarrier + 1-D bulk copy (TMA) ---------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// Shared-memory loads that keep their program order (volatile asm): all 16 warps of the CTA issue their exchange loads
// at the same time, so the LSU returns a warp's j-th load after ~16 j wavefronts; issuing them in the order the
// butterflies first use them lets the FP64 work start before the last load has landed.
__device__ __forceinline__ double lds_f64(const double* p) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(const double2* p) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ float lds_f64(const float* p) {          // (float32 builds of the precision study)
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ float2 lds_v2f64(const float2* p) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(smem_u32(p)));
    return v;
}
// warp-uniform tables in the precision of the pass that reads them
__device__ __forceinline__ double  tab_win(int r, int n, double) { return c_win[r][n]; }
__device__ __forceinline__ double2 tab_tw(int r, int k, double) { return c_tw[r][k]; }
__device__ __forceinline__ double2 tab_wh(int r, int k, double) { return c_wh[r][k]; }
#if STX_K_P1_F32 || STX_K_P2_F32
__device__ __forceinline__ float  tab_win(int r, int n, float) { return c_win_f[r][n]; }
__device__ __forceinline__ float2 tab_tw(int r, int k, float) { return c_tw_f[r][k]; }
__device__ __forceinline__ float2 tab_wh(int r, int k, float) { return c_wh_f[r][k]; }
#endif
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
// The PCM is read exactly once: evict-first in L2, so that the raw log-mel this kernel writes (61 MB on cfg2, re-read by
// k_normalize) stays resident instead of the 123 MB input stream.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// Samples [lo, hi) of the clip that the bulk copy stages for the tile starting at sample s0
// (stage[kLead + i] = x[s0 + i]).  Both ends are multiples of 4 samples so that the copy is 16-byte
// aligned and sized; everything else (the <= 3 tail samples of a clip, or the whole range when the clip
// is not 16-byte aligned in memory) is read with plain loads in the conversion pass.
struct StageRange { int lo, hi; };
__device__ __forceinline__ StageRange stage_range(int s0, int n, bool aligned) {
    StageRange r;
    r.lo = max(s0 - kLead, 0);
    r.hi = aligned ? min(s0 + kTileSamples + 4, n & ~3) : r.lo;
    if (r.hi < r.lo) r.hi = r.lo;
    return r;
}

__constant__ __align__(16) float c_melw[kMelWeights];     // [slot][warp][mel_len(slot)]
__constant__ int c_melfirst[kMel];

// Filter-walking mel slot with warp-uniform weights from the constant bank: only the power spectrum goes through the
// shared-memory pipe (704 reads per frame instead of 704 + 176 + 80)
template <int kSlot>
__device__ __forceinline__ float mel_slot_c(const float* __restrict__ Pl, int warp) {
    constexpr int L = mel_len(kSlot);
    const float* pk = Pl + c_melfirst[16 * kSlot + warp] * kTile;
    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
    for (int q = 0; q < L / 4; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(&c_melw[mel_off(kSlot) + warp * L + 4 * q]);
        acc0 = fmaf(w.x, pk[(4 * q + 0) * kTile], acc0);
        acc1 = fmaf(w.y, pk[(4 * q + 1) * kTile], acc1);
        acc0 = fmaf(w.z, pk[(4 * q + 2) * kTile], acc0);
        acc1 = fmaf(w.w, pk[(4 * q + 3) * kTile], acc1);
    }
    return ln_pos(fmaxf(acc0 + acc1, kMelFloor));
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
    return ln_pos(fmaxf(acc0 + acc1, kMelFloor));
}

// the same filter, accumulation only (the caller applies floor + ln after ALL its filters' loads have been issued)
template <int kHalfSlot>
__device__ __forceinline__ float mel_acc10(const float* __restrict__ Pl, const float* __restrict__ melw,
                                           const int* __restrict__ melfirst, int w8) {
    constexpr int L = mel_len10(kHalfSlot), PAD = mel_pad10(kHalfSlot);
    const float4* w4 = reinterpret_cast<const float4*>(melw + mel_off10(kHalfSlot) + w8 * PAD);
    const float* pk = Pl + melfirst[8 * kHalfSlot + w8] * kTile;
    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
    for (int q = 0; q < PAD / 4; ++q) {
        const float4 w = w4[q];
        if (4 * q + 0 < L) acc0 = fmaf(w.x, pk[(4 * q + 0) * kTile], acc0);
        if (4 * q + 1 < L) acc1 = fmaf(w.y, pk[(4 * q + 1) * kTile], acc1);
        if (4 * q + 2 < L) acc0 = fmaf(w.z, pk[(4 * q + 2) * kTile], acc0);
        if (4 * q + 3 < L) acc1 = fmaf(w.w, pk[(4 * q + 3) * kTile], acc1);
    }
    return acc0 + acc1;
}
template <int kSlot>
__device__ __forceinline__ float mel_acc(const float* __restrict__ Pl, const float* __restrict__ melw,
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
    return acc0 + acc1;
}

template <bool kPeak>
__global__ void __launch_bounds__(kThreads, 1)
k_frames(const float* __restrict__ pcm, const long long* __restrict__ offsets, const int* __restrict__ lengths,
         const float* __restrict__ peaks, const KTables* __restrict__ tab, int T_pad, int chunk_frames,
         int chunks_per_clip, float* __restrict__ out, long long* __restrict__ partials) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const int b = blockIdx.y;
    const int chunk = blockIdx.x;
    const int n = lengths[b];
    const int T = min(n >= kFrame ? 1 + (n - kFrame) / kHop : 0, chunks_per_clip * chunk_frames);
    const int t_begin = chunk * chunk_frames;
    if (t_begin >= T) return;                       // uniform per CTA
    const int t_end = min(T, t_begin + chunk_frames);
    const float* clip = pcm + offsets[b];
    const bool aligned = (reinterpret_cast<unsigned long long>(clip) & 15ull) == 0;
    const float peak = kPeak ? peaks[b] : 1.0f;
    float* out_b = out + (size_t)b * T_pad * kMel;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform: table reads stay on the constant path

    // first tile's PCM: issue the bulk copy before anything else
    if (tid == 0) {
        mbar_init(&sm.mbar, 1);
        mbar_init(&sm.cbar, kWarps);
        const StageRange sr = stage_range(t_begin * kHop, n, aligned);
        if (sr.hi > sr.lo) {
            mbar_expect_tx(&sm.mbar, (unsigned)(sr.hi - sr.lo) * 4u);
            bulk_g2s(sm.stage + (sr.lo - (t_begin * kHop - kLead)), clip + sr.lo, (unsigned)(sr.hi - sr.lo) * 4u, &sm.mbar);
        }
    }
    for (int i = tid; i < kMelWeights; i += kThreads) sm.melw[i] = tab->melw[i];
    if (tid < kMel) sm.melfirst[tid] = tab->melfirst[tid];
    __syncthreads();                                // mbarrier init + tables visible

    float* outstage = reinterpret_cast<float*>(&sm.ex[0][0][0]);
    unsigned long long s1 = 0, s2h = 0, s2l = 0;    // fixed-point statistics of bin tid % 80 over the rows tid / 80 + 6 i
    const int sbin = tid % kMel, srow = tid / kMel;
    unsigned parity = 0, cparity = 0;

    // convert(tc): staged PCM of the tile at frame tc -> d[i] = x[i] - 0.97 x[i-1] (float64), rows of 161.  It runs in the
    // SAME phase as the previous tile's store / statistics (d aliases the power spectrum, which is dead by then), so that its
    // LDS -> F2F -> DFMA -> STS chains overlap the global stores, and a tile costs four CTA barriers instead of five.
    auto convert = [&](const int tc) {
        // ---- convert: staged PCM -> d[i] = x[i] - 0.97 x[i-1] (float64), rows of 161 ----
        const int s0 = tc * kHop;
        const StageRange sr = stage_range(s0, n, aligned);
        if (sr.hi > sr.lo) { mbar_wait(&sm.mbar, parity); parity ^= 1; }
        auto sample = [&](int g) -> float {          // x[g] of this clip, 0 outside
            float v = 0.0f;
            if (g >= sr.lo && g < sr.hi) v = sm.stage[g - s0 + kLead];
            else if (g >= 0 && g < n) v = __ldg(clip + g);
            if (kPeak) v = v / peak;                  // float32 division, like numpy's (R/processor.py:92)
            return v;
        };
        if (sr.lo == s0 - kLead && sr.hi == s0 + kTileSamples + 4) {
            // whole tile staged and inside the clip: no range checks.  480 threads, thread (r0, c0) = (tid / 160, tid % 160)
            // converts samples 160 (r0 + 3 j) + c0: no division in the loop, and the iterations are independent
            // (unrolled: the LDS -> F2F -> DFMA -> STS chains of several samples overlap)
            if (tid < 3 * kHop) {
                const int r0 = tid / kHop, c0 = tid - r0 * kHop;
                const float* src = sm.stage + kLead + r0 * kHop + c0;
                r1_t* dst = sm.u.d + r0 * kDRow + c0;
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                    if (j < 11 || r0 * kHop + c0 < kTileSamples - 33 * kHop) {      // rows 0..32 are whole, row 33 holds 80 samples
                        float xm = src[3 * kHop * j - 1], xi = src[3 * kHop * j];
                        if (kPeak) { xm = xm / peak; xi = xi / peak; }
                        dst[3 * kDRow * j] = codelets::fma_((r1_t)-0.97, (r1_t)xm, (r1_t)xi);
                    }
                }
            }
            if (tid < kTile) {
                float xa = sm.stage[kLead + tid * kHop + kFrame - 1], xz = sm.stage[kLead - 1 + tid * kHop];
                if (kPeak) { xa = xa / peak; xz = xz / peak; }
                sm.xb[tid] = (r1_t)0.97 * ((r1_t)xa - (r1_t)xz);
            }
        } else {
#pragma unroll 1
            for (int i = tid; i < kTileSamples; i += kThreads)
                sm.u.d[i + (unsigned)i / kHop] = codelets::fma_((r1_t)-0.97, (r1_t)sample(s0 + i - 1), (r1_t)sample(s0 + i));
            if (tid < kTile)
                sm.xb[tid] = (r1_t)0.97 * ((r1_t)sample(s0 + tid * kHop + kFrame - 1) - (r1_t)sample(s0 + tid * kHop - 1));
        }
    };
    auto prefetch = [&](const int tn) {             // PCM of the tile at frame tn lands while the tile before it is transformed
        if (tid == 0 && tn < t_end) {
            const int sn = tn * kHop;
            const StageRange nx = stage_range(sn, n, aligned);
            if (nx.hi > nx.lo) {
                mbar_expect_tx(&sm.mbar, (unsigned)(nx.hi - nx.lo) * 4u);
                bulk_g2s(sm.stage + (nx.lo - (sn - kLead)), clip + nx.lo, (unsigned)(nx.hi - nx.lo) * 4u, &sm.mbar);
            }
        }
    };
    // the first trip (t0 = t_begin - kTile) only converts the first tile: ONE copy of every phase in the code
    for (int t0 = t_begin - kTile; t0 < t_end; t0 += kTile) {
      if (t0 >= t_begin) {
        // ---- window + pass 1 (warp = n2) ----
        {
            const r1_t* D = sm.u.d + kDRow * lane + warp;
            r1_t y[25];
            r1_t sa = 0, sb = 0;
            // first-use order of the radix-4 decimation in time: y[q], y[q+16], y[q+8], y[q+24], y[q+4], y[q+20], y[q+12]
            constexpr int kOrder1[25] = {0, 16, 8, 24, 4, 20, 12, 1, 17, 9, 5, 21, 13, 2, 18, 10, 6, 22, 14, 3, 19, 11, 7, 23, 15};
            r1_t v1[25];
#pragma unroll
            for (int j = 0; j < 25; ++j) {
                const int n1 = kOrder1[j];
                v1[n1] = lds_f64(D + 16 * n1 + (n1 >= 10) + (n1 >= 20));
            }
#pragma unroll
            for (int n1 = 0; n1 < 25; ++n1) {
                y[n1] = tab_win(warp, n1, r1_t()) * v1[n1];
                if (n1 & 1) sb += v1[n1]; else sa += v1[n1];
            }
            sm.psum[lane][warp] = sa + sb;
            r1_t re[17], im[17];
            codelets::k_pass1<r1_t>(y, re, im);
            sm.ex0[warp][lane] = (r2_t)re[0];
            sm.ex16[warp][lane] = (r2_t)re[16];
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) {
                const auto t = tab_tw(warp, k1, r1_t());
                r2x2_t o;
                o.x = (r2_t)codelets::fma_(re[k1], t.x, -(im[k1] * t.y));
                o.y = (r2_t)codelets::fma_(re[k1], t.y, im[k1] * t.x);
                sm.ex[k1 - 1][warp][lane] = o;
            }
        }
        __syncthreads();                            // exchange complete; d is dead, its storage becomes the power spectrum

        // ---- c = 0.03 * mean(frame): warp w sums the 16 partials of frames 2 w and 2 w + 1 ----
        {
            const int fr = 2 * warp + (lane >> 4), r = lane & 15;
            r1_t v = sm.psum[fr][r];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            if (r == 0) sm.cval[fr] = (r2_t)((v - sm.xb[fr]) * (r1_t)(1.0 / 400.0));
            if (warp == 1) sm.u.P[0][lane] = 0.0f;   // padded mel filters may touch bin 0 with a zero weight
            // c is only needed after the DFT-16 below: one arrival per warp on an mbarrier now, the wait is there,
            // so that no warp idles here
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sm.cbar)) : "memory");
        }

        // ---- pass 2 (warp = k1), DC correction, power ----
        {
            r2_t c = 0;
            auto put = [&](int bin, r2_t xr, r2_t xi, r2x2_t wh) {
                const r2_t a = codelets::fma_(-c, wh.x, xr), bb = codelets::fma_(-c, wh.y, xi);
                sm.u.P[bin][lane] = power_to_f32(codelets::fma_(a, a, bb * bb));
            };
            if (warp == 0) {
                r2_t a[16], r[16];
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) { a[n2] = sm.ex0[n2][lane]; r[n2] = sm.ex16[n2][lane]; }
                r2_t e0r[7], e0i[7], e16r[8], e16i[8];
                codelets::k_pass2_edge<r2_t>(a, r, e0r, e0i, e16r, e16i);
                mbar_wait(&sm.cbar, cparity);
                c = sm.cval[lane];
#pragma unroll
                for (int k2 = 1; k2 < 8; ++k2) put(32 * k2, e0r[k2 - 1], e0i[k2 - 1], tab_wh(0, k2, r2_t()));
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) put(16 + 32 * k2, e16r[k2], e16i[k2], tab_wh(0, 8 + k2, r2_t()));
            } else {
                r2_t xr[16], xi[16], yr[16], yi[16];
                constexpr int kOrder2[16] = {0, 8, 4, 12, 1, 9, 5, 13, 2, 10, 6, 14, 3, 11, 7, 15};
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n2 = kOrder2[j];
                    const r2x2_t v = lds_v2f64(&sm.ex[warp - 1][n2][lane]);
                    xr[n2] = v.x; xi[n2] = v.y;
                }
                codelets::dft16<r2_t>(xr, xi, yr, yi);
                mbar_wait(&sm.cbar, cparity);
                c = sm.cval[lane];
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2)
                    put(k2 < 8 ? warp + 32 * k2 : 512 - warp - 32 * k2, yr[k2], yi[k2], tab_wh(warp, k2, r2_t()));
            }
            cparity ^= 1;
        }
        __syncthreads();

        // ---- sparse mel + ln (warp w: bins w, w + 16, ..., w + 64; weights are warp-uniform broadcasts) ----
        {
            const float* Pl = &sm.u.P[0][lane];
            float* orow = outstage + lane * kOutRow + warp;
            orow[0]  = mel_slot<0>(Pl, sm.melw, sm.melfirst, warp);
            orow[16] = mel_slot<1>(Pl, sm.melw, sm.melfirst, warp);
            orow[32] = mel_slot<2>(Pl, sm.melw, sm.melfirst, warp);
            orow[48] = mel_slot<3>(Pl, sm.melw, sm.melfirst, warp);
            orow[64] = mel_slot<4>(Pl, sm.melw, sm.melfirst, warp);
        }
        __syncthreads();

        // ---- coalesced store of the tile's rows (one contiguous block of out) + statistics ----
        if (tid < kStatThreads) {
            const int rows = min(t_end - t0, kTile);
            const int keep = min(T_pad - t0, rows);           // frames >= T_pad count for the statistics but are not stored
            float* dst = out_b + (size_t)t0 * kMel + tid;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int row = srow + 6 * i;
                if (row < rows) {
                    const float v = outstage[row * kOutRow + sbin];
                    if (row < keep) dst[i * kStatThreads] = v;
                    const double vd = (double)v;
                    const double sq = vd * vd, hi = sq + kFixH, lo = sq - (hi - kFixH);      // all exact
                    s1 += (unsigned long long)__double_as_longlong(vd + kFix1) - (unsigned long long)__double_as_longlong(kFix1);
                    s2h += (unsigned long long)__double_as_longlong(hi) - (unsigned long long)__double_as_longlong(kFixH);
                    s2l += (unsigned long long)__double_as_longlong(lo + kFixL) - (unsigned long long)__double_as_longlong(kFixL);
                }
            }
        }
        // the next tile's conversion only touches d / xb / stage: the mel stage's reads of the power spectrum (aliased with
        // d) are all before the __syncthreads above, and xb / cval of this tile were consumed in pass 2
      }
        if (t0 + kTile < t_end) convert(t0 + kTile);
        __syncthreads();                            // d ready, staging free again; staged rows read before pass 1 rewrites them
        prefetch(t0 + 2 * kTile);
    }

    // ---- per-chunk statistics: ordered reduction over the 6 row groups ----
    __syncthreads();
    unsigned long long* red = reinterpret_cast<unsigned long long*>(&sm.ex[0][0][0]);     // [3][6][80]
    if (tid < kStatThreads) {
        red[tid] = s1;
        red[kStatThreads + tid] = s2h;
        red[2 * kStatThreads + tid] = s2l;
    }
    __syncthreads();
    if (tid < kStatWords) {
        const int which = tid / kMel, m = tid - which * kMel;
        unsigned long long acc = 0;
#pragma unroll
        for (int g = 0; g < 6; ++g) acc += red[which * kStatThreads + g * kMel + m];
        partials[((size_t)b * chunks_per_clip + chunk) * kStatWords + tid] = (long long)acc;
    }
}

// =================================================================================================================
// k_frames_duo: the same arithmetic as k_frames, organised as TWO INDEPENDENT HALF-CTAs of 8 warps, each on its own
// 32-frame tile, so that one group's shared-memory phases overlap the other's FP64 phases (in k_frames all 16 warps move
// through the phases together: FP64 pipe 35 % busy + LSU pipe 48 % busy, never at the same time; a synthetic FFT core,
// tools/microbench_overlap.cu, runs 1.6 x faster with two staggered groups).
//
// Two tiles in flight need two exchange buffers, and a 32-frame float64 exchange is 128 KB.  Each group therefore exchanges
// in TWO HALVES through a 64 KB buffer: pass 1 (two roles per warp) stores the twiddled rows k1 = 0..7 and 16 and STASHES
// rows 8..15 in tensor memory (tcgen05.st: a warp's own 32 TMEM lanes, 32 columns per role -- TMEM is used as a private
// 256 KB register file extension here, no MMA involved); after pass 2 has read the first half, the stash comes back
// (tcgen05.ld) into the same buffer for the second half.  Between a tile's last exchange load and the next tile's first
// store the buffer doubles as the TMA landing zone of the next tile's PCM and as the staging area of the log-mel rows.
//
// Group-private barriers (bar.sync 1 + g, 256).  Per tile: d ready | H1 stored | H1 loaded | H2 stored | P complete | rows staged.
// Measured on cfg2: 224 us against 244 us for k_frames (FP64 pipe 39 %, LSU 53 %: the two groups drift apart on their own;
// delaying group 1 by a fraction of the tile period only adds the delay).  STX_K_SINGLE=1 selects k_frames for A/B runs.
constexpr int kGWarps = 8;
constexpr int kGThreads = kGWarps * 32;          // 256
constexpr int kGStat = 3 * kMel;                 // 240 threads of a group store 3 rows of 80 bins per step
constexpr int kStageBytes = kStage * 4;          // 21 472
constexpr int kOutStageOff = 24576;              // byte offset of the staged log-mel rows inside the exchange buffer
constexpr int kRedOff = 40960;                   // byte offset of the per-item statistics reduction inside the exchange buffer
static_assert(kStageBytes <= kOutStageOff && kOutStageOff + kTile * kOutRow * 4 <= 65536, "landing zone and staged rows share the exchange buffer");

// One 32-frame tile of a work item (clip b, chunk): everything the phases need, kept in SHARED memory and read where it is
// used, so that no per-tile state is live in registers across the FFT phases
struct TileDesc {
    const float* clip; float* out_b; long long* part;
    int n, t0, t_end, item;
    float peak;
    int aligned, valid, last;    // last: the item ends with this tile (flush the statistics)
    int b, nitems;               // clip index and the number of work items of that clip
};
struct SmemG {
    double2 ex[8][16][kTile];    // one HALF of the exchange: [slot][n2][lane].  H1: slot 0 = (row 0, row 16) (both real), slots
                                 // 1..7 = rows 1..7;  H2: slot s = row 8 + s
    union {
        double d[kDBuf];
        float  P[256][kTile];
    } u;
    double  psum[kTile][17];
    double  cval[kTile];
    double  xb[kTile];
    unsigned long long mbar;     // TMA completion
    unsigned long long cbar;     // cval ready (one arrival per warp of the group)
    TileDesc desc[2];            // the tile in flight and the next one (written by the group's thread 0)
    int is_last;                 // this group completed the last work item of a clip (it reduces the clip's statistics)
#if STX_K_TRACE == 1
    unsigned tlast, tacc[8];
#elif STX_K_TRACE == 2
    unsigned tl[8], tw[8][8];
#endif
};
struct SmemDuo {
    SmemG g[2];
    float melw[kMelWeights];
    int   melfirst[kMel];
    unsigned tmem_base;
};
static_assert(sizeof(SmemDuo) <= 227 * 1024, "one CTA per SM must fit in 227 KB");
constexpr unsigned kTmemCols = 256;              // 4 warps per TMEM lane quarter x 2 roles x 32 columns

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(kGThreads) : "memory"); }

// 16 doubles <-> 32 TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_st16(unsigned taddr, const double (&v)[16]) {
    unsigned r[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) { r[2 * i] = (unsigned)__double2loint(v[i]); r[2 * i + 1] = (unsigned)__double2hiint(v[i]); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                    "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                    "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
// the same stash, one double (= one natural register pair) per instruction: no moves to marshal 32 consecutive registers
__device__ __forceinline__ void tmem_st16_pairs(unsigned taddr, const double (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};"
                     :: "r"(taddr + 2u * i), "r"((unsigned)__double2loint(v[i])), "r"((unsigned)__double2hiint(v[i])) : "memory");
}
__device__ __forceinline__ void tmem_st16_quads(unsigned taddr, const double (&v)[16]) {      // (one complex value per instruction)
#pragma unroll
    for (int i = 0; i < 8; ++i)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                     :: "r"(taddr + 4u * i), "r"((unsigned)__double2loint(v[2 * i])), "r"((unsigned)__double2hiint(v[2 * i])),
                        "r"((unsigned)__double2loint(v[2 * i + 1])), "r"((unsigned)__double2hiint(v[2 * i + 1])) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, double (&v)[16]) {
    unsigned r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}

// mean and 1/sqrt(var + 1e-7) of one mel bin from the clip's fixed-point sums (sum x on the 2^-32 grid, sum x^2 as a 2^-20-grid
// part plus a 2^-56-grid remainder); var with ddof = 1 (…seamless_m4t.py:257-262).  The mean is returned as two float32
// words (k_normalize subtracts both).  A single frame has no ddof=1 variance: numpy returns NaN there and so do we.
__device__ __forceinline__ void clip_stats(long long s1, long long s2h, long long s2l, int T_all, float& mean_hi, float& mean_lo,
                                           float& rstd) {
    const double a1 = (double)s1 * (1.0 / 4294967296.0);
    const double a2 = (double)s2h * (1.0 / 1048576.0) + (double)s2l * (1.0 / 72057594037927936.0);
    const double mean = a1 / (double)T_all;
    double var = T_all > 1 ? (a2 - a1 * mean) / (double)(T_all - 1) : __longlong_as_double(0x7ff8000000000000LL);
    if (var < 0.0) var = 0.0;                                  // rounding of a constant column; keeps NaN
    mean_hi = (float)mean;
    mean_lo = (float)(mean - (double)(float)mean);
    rstd = (float)(1.0 / sqrt(var + 1e-7));
}

template <bool kPeak>
__global__ void __launch_bounds__(kThreads, 1)
k_frames_duo(const float* __restrict__ pcm, const long long* __restrict__ offsets, const int* __restrict__ lengths,
             const float* __restrict__ peaks, const KTables* __restrict__ tab, int B, int T_pad, int chunk_frames,
             int chunks_per_clip, float* __restrict__ out, long long* __restrict__ partials, const int* __restrict__ sched,
             unsigned* __restrict__ done, float* __restrict__ cstat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemDuo& sm = *reinterpret_cast<SmemDuo*>(smem_raw);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = warp >> 3;                        // group
    const int w8 = warp & 7;                        // warp within the group: roles w8 and w8 + 8, row tasks w8 and 8 + w8
    const int tl = tid & (kGThreads - 1);           // thread within the group
    SmemG& sg = sm.g[g];
    float* stage = reinterpret_cast<float*>(&sg.ex[0][0][0]);                         // TMA landing zone (between tiles)
    float* outstage = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(&sg.ex[0][0][0]) + kOutStageOff);
    unsigned long long* red = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(&sg.ex[0][0][0]) + kRedOff);

    // PERSISTENT: the grid is one CTA per SM and every GROUP streams through its own list of work items, item =
    // (clip b, chunk of chunk_frames frames) from the schedule, positions G, G + groups, ... with G = 2 blockIdx.x + g.  The TMEM
    // allocation, the tables and the pipeline fill are paid once per kernel instead of once per chunk (2.3 us each, 7 % of
    // the non-persistent form), and the groups never meet again after the first barrier, so they drift into complementary
    // phases on their own.
#if STX_K_SOLO
    const int item_step = gridDim.x;                // (experiment) group 0 does all the work, group 1 idles: phase times without a neighbour
#else
    const int item_step = 2 * gridDim.x;
#endif
    // (thread 0 of the group only) first tile of the work item at position `pos` of the schedule k_schedule wrote:
    // sched[0] = number of NON-EMPTY items, sched[1 + pos] = b * chunks_per_clip + chunk.  Group G takes positions G,
    // G + groups, ...: every group gets the same number of non-empty items (+- 1) however ragged the batch is
    // Without a schedule (sched == nullptr: the caller vouches that all clips have the same length, so every item is
    // non-empty and the plain round-robin over all items is already balanced) positions are item numbers.
    auto open_item = [&](int pos, TileDesc& d) {
        const int n_pos = sched ? __ldg(sched) : B * chunks_per_clip;
        d.valid = 0; d.last = 0; d.item = pos;
        for (; pos < n_pos; pos += item_step) {
            const int item = sched ? __ldg(sched + 1 + pos) : pos;
            const int b = item / chunks_per_clip, chunk = item - b * chunks_per_clip;
            const int n = __ldg(lengths + b);
            // (a clip longer than the max_length the caller sized the call for is processed up to that length)
            const int T = min(n >= kFrame ? 1 + (n - kFrame) / kHop : 0, chunks_per_clip * chunk_frames);
            const int t_begin = chunk * chunk_frames;
            if (t_begin >= T) continue;              // (only without a schedule, and only if the caller's promise was wrong)
            d.valid = 1; d.item = pos; d.n = n; d.t0 = t_begin; d.t_end = min(T, t_begin + chunk_frames);
            d.clip = pcm + __ldg(offsets + b);
            d.aligned = (reinterpret_cast<unsigned long long>(d.clip) & 15ull) == 0;
            d.peak = kPeak ? __ldg(peaks + b) : 1.0f;
            d.out_b = out + (size_t)b * T_pad * kMel;
            d.part = partials + (size_t)item * kStatWords;
            d.b = b; d.nitems = (T + chunk_frames - 1) / chunk_frames;
            break;
        }
    };
    // (thread 0 of the group only) d = the tile after c; marks c as the last tile of its item when the item changes
    auto next_tile = [&](TileDesc& c, TileDesc& d) {
        if (c.t0 + kTile < c.t_end) { d = c; d.t0 += kTile; c.last = 0; }
        else { open_item(c.item + item_step, d); c.last = 1; }
    };

    auto prefetch = [&](const TileDesc& d) {        // (thread 0 of the group) PCM of the tile -> the (idle) exchange buffer
        if (d.valid) {
            const int sn = d.t0 * kHop;
            const StageRange nx = stage_range(sn, d.n, d.aligned);
            if (nx.hi > nx.lo) {
                mbar_expect_tx(&sg.mbar, (unsigned)(nx.hi - nx.lo) * 4u);
                bulk_g2s(stage + (nx.lo - (sn - kLead)), d.clip + nx.lo, (unsigned)(nx.hi - nx.lo) * 4u, &sg.mbar);
            }
        }
    };
    if (tl == 0) {
        mbar_init(&sg.mbar, 1);
        mbar_init(&sg.cbar, kGWarps);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
#if STX_K_MEL_GEN == 0 && STX_K_MEL_EXACT
    for (int i = tid; i < kMelWeights; i += kThreads) sm.melw[i] = tab->melw10[i];
    if (tid < kMel) sm.melfirst[tid] = tab->melfirst10[tid];
#elif STX_K_MEL_GEN == 0
    for (int i = tid; i < kMelWeights; i += kThreads) sm.melw[i] = tab->melw[i];
    if (tid < kMel) sm.melfirst[tid] = tab->melfirst[tid];
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                // mbarrier init, tables, TMEM base visible
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");      // the schedule (k_schedule, the programmatic primary, if any) is complete
#if STX_K_NORM_PDL
    // k_normalize (the programmatic dependent) may be scheduled onto SMs as they drain; its CTAs block in griddepcontrol.wait
    // until this whole grid has completed and flushed, so only its launch latency is hidden
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    if (tl == 0) {
#if STX_K_SOLO
        if (g == 0) open_item(blockIdx.x, sg.desc[0]); else sg.desc[0].valid = 0;
#else
        open_item(2 * blockIdx.x + g, sg.desc[0]);
#endif
        prefetch(sg.desc[0]);
    }
    group_bar(g);                                   // the first descriptor is visible
    // this thread's stash: TMEM lanes of the warp's quarter, 32 columns per role
    const unsigned tstash = sm.tmem_base + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)((warp >> 2) * 64);

    unsigned long long s1 = 0, s2h = 0, s2l = 0;    // fixed-point statistics of bin tl % 80 over the rows tl / 80 + 3 i
    const int sbin = tl % kMel, srow = tl / kMel;
    unsigned parity = 0, cparity = 0;
#if STX_K_TRACE == 1
    if (tl < 8) sg.tacc[tl] = 0u;
    if (tl == 0) sg.tlast = clock_lo();
#elif STX_K_TRACE == 2
    if (tl < 64) sg.tw[tl >> 3][tl & 7] = 0u;
    if (tl < 8) sg.tl[tl] = clock_lo();
    group_bar(g);
#endif

    // convert(td): landed PCM of tile td -> d[i] = x[i] - 0.97 x[i-1] (float64), rows of 161
    auto convert = [&](const TileDesc& td) {
        const int n = td.n;
        const float* clip = td.clip;
        const bool aligned = td.aligned != 0;
        const float peak = td.peak;
        const int s0 = td.t0 * kHop;
        const StageRange sr = stage_range(s0, n, aligned);
        if (sr.hi > sr.lo) { mbar_wait(&sg.mbar, parity); parity ^= 1; }
        auto sample = [&](int gi) -> float {        // x[gi] of this clip, 0 outside
            float v = 0.0f;
            if (gi >= sr.lo && gi < sr.hi) v = stage[gi - s0 + kLead];
            else if (gi >= 0 && gi < n) v = __ldg(clip + gi);
            if (kPeak) v = v / peak;                  // float32 division, like numpy's (R/processor.py:92)
            return v;
        };
        if (sr.lo == s0 - kLead && sr.hi == s0 + kTileSamples + 4) {
#if (STX_K_ILP & 4)
            // whole tile landed and inside the clip.  215 threads, thread t converts the RUN of 25 consecutive samples
            // 25 t .. 25 t + 24: x[i - 1] of one sample is x[i] of the one before, so 26 conversions (quarter-rate XU
            // instructions) give 25 outputs instead of 50, and 26 loads instead of 50.  The odd stride keeps both the
            // 4-byte loads and the 8-byte stores free of bank conflicts; a run crosses at most one row of d (160 + 1 pad).
            constexpr int kRun = 25, kRunThreads = (kTileSamples + kRun - 1) / kRun;      // 215
            if (tl < kRunThreads) {
                const int i0 = kRun * tl;
                const int q0 = (5 * tl) >> 5;               // = i0 / 160
                const int bnd = kHop * (q0 + 1) - i0;             // first j of the run that lies in the next row
                const float* src = stage + kLead + i0;
                double* dst = sg.u.d + i0 + q0;
                float xf[kRun + 1];
#pragma unroll
                for (int j = 0; j <= kRun; ++j) xf[j] = src[j - 1];    // (the last run reads up to 15 floats past the tile, inside the buffer)
                if (kPeak) {
#pragma unroll
                    for (int j = 0; j <= kRun; ++j) xf[j] = xf[j] / peak;
                }
                double xd[kRun + 1];
#pragma unroll
                for (int j = 0; j <= kRun; ++j) xd[j] = (double)xf[j];
                // (the last run writes 15 values past the tile's samples: inside d, never read)
#pragma unroll
                for (int j = 0; j < kRun; ++j) dst[j + (j >= bnd ? 1 : 0)] = fma(-0.97, xd[j], xd[j + 1]);
            }
#elif STX_K_CVT_ROWS
            // whole tile landed and inside the clip.  160 threads, one column of the 34 rows of d each: both addresses are
            // base + constant * row, so an iteration is 2 LDS + 2 F2F + DFMA + STS and nothing else (the half-row form below
            // spends as many integer instructions on its addresses as on the conversion); the last warp, otherwise idle in
            // this phase, computes xb
            if (tl < kHop) {
                const float* src = stage + kLead + tl;
                double* dst = sg.u.d + tl;
#pragma unroll
                for (int r = 0; r < 34; ++r) {
                    if (r < 33 || tl < kTileSamples - 33 * kHop) {
                        float xm = src[kHop * r - 1], xi = src[kHop * r];
                        if (kPeak) { xm = xm / peak; xi = xi / peak; }
                        dst[kDRow * r] = fma(-0.97, (double)xm, (double)xi);
                    }
                }
            }
            if (tl >= kGThreads - kTile) {
                const int f = tl - (kGThreads - kTile);
                float xa = stage[kLead + f * kHop + kFrame - 1], xz = stage[kLead - 1 + f * kHop];
                if (kPeak) { xa = xa / peak; xz = xz / peak; }
                sg.xb[f] = 0.97 * ((double)xa - (double)xz);
            }
#else
            // whole tile landed and inside the clip.  240 threads, thread (u, c) = (tl / 80, tl % 80) converts the half rows
            // u + 3 j (67 half rows of 80 samples): no division in the loop, independent iterations
#if STX_K_CVT_LINEAR
            // ... with every address a per-thread base plus a compile-time constant: half row hr = u + 3 j sits at
            // d[80 hr + (hr >> 1) + c], and (u + 3 j) >> 1 is linear in j once even and odd j are taken apart
            // (j = 2 m: 3 m + (u >> 1); j = 2 m + 1: 3 m + ((u + 3) >> 1)), so the loop body is 2 LDS + 2 F2F + DFMA + STS
            if (tl < kGStat) {
                const int u = tl / 80, c = tl - u * 80;
                const float* src = stage + kLead + 80 * u + c;
                double* de = sg.u.d + 80 * u + c + (u >> 1);                       // j even
                double* dq = sg.u.d + 80 * (u + 3) + c + ((u + 3) >> 1);           // j odd
#pragma unroll
                for (int m = 0; m < 12; ++m) {
                    {
                        float xm = src[480 * m - 1], xi = src[480 * m];
                        if (kPeak) { xm = xm / peak; xi = xi / peak; }
                        if (m < 11 || u == 0) de[483 * m] = fma(-0.97, (double)xm, (double)xi);      // (j = 22: half row 66 only)
                    }
                    if (m < 11) {
                        float xm = src[480 * m + 240 - 1], xi = src[480 * m + 240];
                        if (kPeak) { xm = xm / peak; xi = xi / peak; }
                        dq[483 * m] = fma(-0.97, (double)xm, (double)xi);
                    }
                }
            }
#else
            if (tl < kGStat) {
                const int u = tl / 80, c = tl - u * 80;
#pragma unroll
                for (int j = 0; j < 23; ++j) {
                    const int hr = u + 3 * j;
                    if (hr < 67) {
                        const float* src = stage + kLead + 80 * hr + c;
                        float xm = src[-1], xi = src[0];
                        if (kPeak) { xm = xm / peak; xi = xi / peak; }
#if STX_K_CVT_SPLIT == 2
                        sg.u.d[80 * hr + c + (hr >> 1)] = fma(-0.97, f32_to_f64_int(xm), f32_to_f64_int(xi));
#elif STX_K_CVT_SPLIT
                        sg.u.d[80 * hr + c + (hr >> 1)] = fma(-0.97, f32_to_f64_int(xm), (double)xi);
#else
                        sg.u.d[80 * hr + c + (hr >> 1)] = fma(-0.97, (double)xm, (double)xi);
#endif
                    }
                }
            }
#endif
#endif
#if !STX_K_CVT_ROWS || (STX_K_ILP & 4)
            if (tl < kTile) {
                float xa = stage[kLead + tl * kHop + kFrame - 1], xz = stage[kLead - 1 + tl * kHop];
                if (kPeak) { xa = xa / peak; xz = xz / peak; }
                sg.xb[tl] = 0.97 * ((double)xa - (double)xz);
            }
#endif
        } else {
#pragma unroll 1
            for (int i = tl; i < kTileSamples; i += kGThreads)
                sg.u.d[i + (unsigned)i / kHop] = fma(-0.97, (double)sample(s0 + i - 1), (double)sample(s0 + i));
            if (tl < kTile)
                sg.xb[tl] = 0.97 * ((double)sample(s0 + tl * kHop + kFrame - 1) - (double)sample(s0 + tl * kHop - 1));
        }
    };

    // the first trip only converts the group's first tile: ONE copy of every phase in the code
    int slot = 0;
    bool first_trip = true;
    while (sg.desc[first_trip ? 0 : slot].valid) {
      if (!first_trip) {
        if (tl == 0) next_tile(sg.desc[slot], sg.desc[slot ^ 1]);      // visible to the group after the next barrier
        // ---- window + pass 1, two roles per warp: rows 0..7 and 16 -> exchange (H1), rows 8..15 -> TMEM stash ----
        auto pass1 = [&](const int role, const int which) {
            const double* D = sg.u.d + kDRow * lane + role;
            double y[25];
            double sa = 0.0, sb = 0.0;
            constexpr int kOrder1[25] = {0, 16, 8, 24, 4, 20, 12, 1, 17, 9, 5, 21, 13, 2, 18, 10, 6, 22, 14, 3, 19, 11, 7, 23, 15};
            double v1[25];
#pragma unroll
            for (int j = 0; j < 25; ++j) {
                const int n1 = kOrder1[j];
                v1[n1] = lds_f64(D + 16 * n1 + (n1 >= 10) + (n1 >= 20));
            }
#pragma unroll
            for (int n1 = 0; n1 < 25; ++n1) {
                y[n1] = c_win[role][n1] * v1[n1];
                if (n1 & 1) sb += v1[n1]; else sa += v1[n1];
            }
            sg.psum[lane][role] = sa + sb;
            double re[17], im[17];
            codelets::k_pass1<double>(y, re, im);
            sg.ex[0][role][lane] = make_double2(re[0], re[16]);
#pragma unroll
            for (int k1 = 1; k1 < 8; ++k1) {
                const double2 t = *reinterpret_cast<const double2*>(&c_tw[role][k1]);
                sg.ex[k1][role][lane] = make_double2(fma(re[k1], t.x, -(im[k1] * t.y)), fma(re[k1], t.y, im[k1] * t.x));
            }
            double hs[16];
#pragma unroll
            for (int k1 = 8; k1 < 16; ++k1) {
                const double2 t = *reinterpret_cast<const double2*>(&c_tw[role][k1]);
                hs[2 * (k1 - 8)] = fma(re[k1], t.x, -(im[k1] * t.y));
                hs[2 * (k1 - 8) + 1] = fma(re[k1], t.y, im[k1] * t.x);
            }
#if STX_K_STASH_PAIRS == 2
            tmem_st16_quads(tstash + (unsigned)(which * 32), hs);
#elif STX_K_STASH_PAIRS
            tmem_st16_pairs(tstash + (unsigned)(which * 32), hs);
#else
            tmem_st16(tstash + (unsigned)(which * 32), hs);
#endif
        };
#if STX_K_H2_TMEM
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");     // the quarter mates have read the previous tile's stash
#endif
        pass1(w8, 0);
        pass1(w8 + kGWarps, 1);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#if STX_K_H2_TMEM
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");    // the stash is read by the quarter mates behind this barrier
#endif
        KTRACE_PRE(0);
        group_bar(g);                               // H1 complete, psum complete; d is dead, its storage becomes the power spectrum
        KTRACE(0);
#if STX_K_H2_TMEM
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#endif

        // ---- c = 0.03 * mean(frame): warp w8 sums the 16 partials of frames 4 w8 .. 4 w8 + 3 ----
        {
            const int fr = 4 * w8 + (lane >> 3), r = lane & 7;
            double v = sg.psum[fr][r] + sg.psum[fr][r + 8];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            if (r == 0) sg.cval[fr] = (v - sg.xb[fr]) * (1.0 / 400.0);
#if STX_K_MEL_GEN != 1
            if (w8 == 1) sg.u.P[0][lane] = 0.0f;     // padded mel filters may touch bin 0 with a zero weight
#endif
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sg.cbar)) : "memory");
        }

        // ---- pass 2, first half: row task w8 (task 0 = the two real rows 0 and 16, task k = row k) ----
        double xr[16], xi[16];
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) {
            const double2 v = lds_v2f64(&sg.ex[w8][n2][lane]);
            xr[n2] = v.x; xi[n2] = v.y;
        }
        KTRACE_PRE(1);
        group_bar(g);                               // H1 is in registers: the exchange buffer is free for H2
        KTRACE(1);
        // ---- the stash comes back: rows 8..15 of both roles -> exchange (H2) ----
        {
            double hs[16];
            tmem_ld16(tstash, hs);
#if STX_K_H2_TMEM
            // rows 8 + s with s = w8 (mod 4) go to a warp of this lane quarter, which reads them from the stash itself
#pragma unroll
            for (int s = 0; s < 8; ++s) if ((s & 3) != (w8 & 3)) sg.ex[s][w8][lane] = make_double2(hs[2 * s], hs[2 * s + 1]);
            tmem_ld16(tstash + 32u, hs);
#pragma unroll
            for (int s = 0; s < 8; ++s) if ((s & 3) != (w8 & 3)) sg.ex[s][w8 + kGWarps][lane] = make_double2(hs[2 * s], hs[2 * s + 1]);
#else
#pragma unroll
            for (int s = 0; s < 8; ++s) sg.ex[s][w8][lane] = make_double2(hs[2 * s], hs[2 * s + 1]);
            tmem_ld16(tstash + 32u, hs);
#pragma unroll
            for (int s = 0; s < 8; ++s) sg.ex[s][w8 + kGWarps][lane] = make_double2(hs[2 * s], hs[2 * s + 1]);
#endif
        }
#if !STX_K_BAR2_LATE
        KTRACE_PRE(2);
        group_bar(g);                               // H2 complete
        KTRACE(2);
#endif

        double c = 0.0;
        auto put = [&](int bin, double pr, double pi, double2 wh) {
            const double a = fma(-c, wh.x, pr), bb = fma(-c, wh.y, pi);
#if STX_K_POWER_F32
            const float af = (float)a, bf = (float)bb;
            sg.u.P[bin][lane] = fmaf(af, af, bf * bf);
#else
            sg.u.P[bin][lane] = power_to_f32(fma(a, a, bb * bb));
#endif
        };
        if (w8 == 0) {
            double e0r[7], e0i[7], e16r[8], e16i[8];
            codelets::k_pass2_edge<double>(xr, xi, e0r, e0i, e16r, e16i);
            mbar_wait(&sg.cbar, cparity);
            c = sg.cval[lane];
#pragma unroll
            for (int k2 = 1; k2 < 8; ++k2) put(32 * k2, e0r[k2 - 1], e0i[k2 - 1], c_wh[0][k2]);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) put(16 + 32 * k2, e16r[k2], e16i[k2], c_wh[0][8 + k2]);
        } else {
            double yr[16], yi[16];
            codelets::dft16<double>(xr, xi, yr, yi);
            mbar_wait(&sg.cbar, cparity);
            c = sg.cval[lane];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2)
                put(k2 < 8 ? w8 + 32 * k2 : 512 - w8 - 32 * k2, yr[k2], yi[k2], c_wh[w8][k2]);
        }
        cparity ^= 1;
#if STX_K_BAR2_LATE
        group_bar(g);                               // H2 complete (its stores drained under the first half's arithmetic)
#endif
        // ---- pass 2, second half: row 8 + w8 ----
        {
            const int row = kGWarps + w8;
            constexpr int kOrder2[16] = {0, 8, 4, 12, 1, 9, 5, 13, 2, 10, 6, 14, 3, 11, 7, 15};
#if STX_K_H2_TMEM
            // roles n2 = q, q + 4, q + 8, q + 12 (q = w8 mod 4) were produced by the two warps of this lane quarter (p = q and
            // q + 4, first and second role each): row 8 + w8 = doubles 2 w8, 2 w8 + 1 of their stashes.  One code path per q,
            // so that the register of every n2 is fixed
            const unsigned tq = sm.tmem_base + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)(g * 128) + 4u * (unsigned)w8;
            auto load_half = [&](auto qc) {
                constexpr int q = decltype(qc)::value;
                unsigned r[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {                // role n2 = q + 4 i: producer warp n2 % 8 of the group, its role n2 / 8
                    const int n2 = q + 4 * i;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(r[i][0]), "=r"(r[i][1]), "=r"(r[i][2]), "=r"(r[i][3])
                                 : "r"(tq + (unsigned)(((n2 & 7) >> 2) * 64 + (n2 >> 3) * 32)) : "memory");
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n2 = kOrder2[j];
                    if ((n2 & 3) == q) continue;
                    const double2 v = lds_v2f64(&sg.ex[w8][n2][lane]);
                    xr[n2] = v.x; xi[n2] = v.y;
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    xr[q + 4 * i] = __hiloint2double((int)r[i][1], (int)r[i][0]);
                    xi[q + 4 * i] = __hiloint2double((int)r[i][3], (int)r[i][2]);
                }
            };
            switch (w8 & 3) {
                case 0: load_half(std::integral_constant<int, 0>()); break;
                case 1: load_half(std::integral_constant<int, 1>()); break;
                case 2: load_half(std::integral_constant<int, 2>()); break;
                default: load_half(std::integral_constant<int, 3>()); break;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");     // (the next pass 1 of the quarter mates overwrites the stash)
#else
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n2 = kOrder2[j];
                const double2 v = lds_v2f64(&sg.ex[w8][n2][lane]);
                xr[n2] = v.x; xi[n2] = v.y;
            }
#endif
            double yr[16], yi[16];
            codelets::dft16<double>(xr, xi, yr, yi);
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2)
                put(k2 < 8 ? row + 32 * k2 : 512 - row - 32 * k2, yr[k2], yi[k2], c_wh[row][k2]);
        }
        KTRACE_PRE(3);
        group_bar(g);                               // power spectrum complete; the exchange buffer is idle until the next pass 1
        KTRACE(3);
        if (tl == 0) prefetch(sg.desc[slot ^ 1]);   // ... and takes the next tile's PCM meanwhile

        // ---- sparse mel + ln ----
        {
            const float* Pl = &sg.u.P[0][lane];
            float* orow = outstage + lane * kOutRow;
#if STX_K_MEL_GEN == 1
            // warp w8 owns a contiguous range of filters and reads every bin they touch ONCE; weights are immediates
            // (generated: mel_k.cuh).  281 shared-memory reads per frame instead of 704 + 176 + 80.
            melk::mel_rows_dispatch(w8, Pl, orow, [](float a) { return ln_pos(fmaxf(a, kMelFloor)); });
#elif STX_K_MEL_GEN == 2
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int wr = w8 + h * kGWarps;
                orow[wr]      = mel_slot_c<0>(Pl, wr);
                orow[wr + 16] = mel_slot_c<1>(Pl, wr);
                orow[wr + 32] = mel_slot_c<2>(Pl, wr);
                orow[wr + 48] = mel_slot_c<3>(Pl, wr);
                orow[wr + 64] = mel_slot_c<4>(Pl, wr);
            }
#elif (STX_K_ILP & 1)
            // all ten filters of the warp are accumulated in registers before the first result is stored: ten independent
            // load -> FMA chains instead of ten serial ones
            float r[10];
#if STX_K_MEL_EXACT
            // mel bin w8 + 8 h + 16 s = bin w8 of half slot 2 s + h
            r[0] = mel_acc10<0>(Pl, sm.melw, sm.melfirst, w8);
            r[1] = mel_acc10<2>(Pl, sm.melw, sm.melfirst, w8);
            r[2] = mel_acc10<4>(Pl, sm.melw, sm.melfirst, w8);
            r[3] = mel_acc10<6>(Pl, sm.melw, sm.melfirst, w8);
            r[4] = mel_acc10<8>(Pl, sm.melw, sm.melfirst, w8);
            r[5] = mel_acc10<1>(Pl, sm.melw, sm.melfirst, w8);
            r[6] = mel_acc10<3>(Pl, sm.melw, sm.melfirst, w8);
            r[7] = mel_acc10<5>(Pl, sm.melw, sm.melfirst, w8);
            r[8] = mel_acc10<7>(Pl, sm.melw, sm.melfirst, w8);
            r[9] = mel_acc10<9>(Pl, sm.melw, sm.melfirst, w8);
#else
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int wr = w8 + h * kGWarps;
                r[5 * h + 0] = mel_acc<0>(Pl, sm.melw, sm.melfirst, wr);
                r[5 * h + 1] = mel_acc<1>(Pl, sm.melw, sm.melfirst, wr);
                r[5 * h + 2] = mel_acc<2>(Pl, sm.melw, sm.melfirst, wr);
                r[5 * h + 3] = mel_acc<3>(Pl, sm.melw, sm.melfirst, wr);
                r[5 * h + 4] = mel_acc<4>(Pl, sm.melw, sm.melfirst, wr);
            }
#endif
#pragma unroll
            for (int i = 0; i < 10; ++i) r[i] = ln_pos(fmaxf(r[i], kMelFloor));
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int sl = 0; sl < 5; ++sl) orow[w8 + h * kGWarps + 16 * sl] = r[5 * h + sl];
#else
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int wr = w8 + h * kGWarps;
                orow[wr]      = mel_slot<0>(Pl, sm.melw, sm.melfirst, wr);
                orow[wr + 16] = mel_slot<1>(Pl, sm.melw, sm.melfirst, wr);
                orow[wr + 32] = mel_slot<2>(Pl, sm.melw, sm.melfirst, wr);
                orow[wr + 48] = mel_slot<3>(Pl, sm.melw, sm.melfirst, wr);
                orow[wr + 64] = mel_slot<4>(Pl, sm.melw, sm.melfirst, wr);
            }
#endif
        }
        KTRACE_PRE(4);
        group_bar(g);
        KTRACE(4);

        // ---- coalesced store of the tile's rows + statistics ----
        const TileDesc& cur = sg.desc[slot];
        const int t0 = cur.t0, t_end = cur.t_end;
        float* const out_b = cur.out_b;
        if (tl < kGStat) {
            const int rows = min(t_end - t0, kTile);
            const int keep = min(T_pad - t0, rows);           // frames >= T_pad count for the statistics but are not stored
            float* dst = out_b + (size_t)t0 * kMel + tl;
#if (STX_K_ILP & 2)
            if (keep == kTile) {
                // full tile, all of it stored: no per-row branches, so the 11 load -> convert -> split -> add chains overlap.
                // Rows 30 and 31 exist for srow < 2 only: the third row group adds an exact zero instead.
                float v[11];
#pragma unroll
                for (int i = 0; i < 10; ++i) v[i] = outstage[(srow + 3 * i) * kOutRow + sbin];
                v[10] = srow < 2 ? outstage[(srow + 30) * kOutRow + sbin] : 0.0f;
#pragma unroll
                for (int i = 0; i < 10; ++i) dst[i * kGStat] = v[i];
                if (srow < 2) dst[10 * kGStat] = v[10];
#if STX_K_STATS_TILE
                double a1[2] = {0.0, 0.0}, a2[2] = {0.0, 0.0};
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                    const double vd = (double)v[i];
                    a1[i & 1] += vd;
                    a2[i & 1] = fma(vd, vd, a2[i & 1]);
                }
                {
                    const double t1 = a1[0] + a1[1], t2 = a2[0] + a2[1];
                    const double hi = t2 + kFixH, lo = t2 - (hi - kFixH);
                    s1 += (unsigned long long)__double_as_longlong(t1 + kFix1) - (unsigned long long)__double_as_longlong(kFix1);
                    s2h += (unsigned long long)__double_as_longlong(hi) - (unsigned long long)__double_as_longlong(kFixH);
                    s2l += (unsigned long long)__double_as_longlong(lo + kFixL) - (unsigned long long)__double_as_longlong(kFixL);
                }
#else
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                    const double vd = (double)v[i];
                    const double sq = vd * vd, hi = sq + kFixH, lo = sq - (hi - kFixH);      // all exact
                    s1 += (unsigned long long)__double_as_longlong(vd + kFix1) - (unsigned long long)__double_as_longlong(kFix1);
                    s2h += (unsigned long long)__double_as_longlong(hi) - (unsigned long long)__double_as_longlong(kFixH);
                    s2l += (unsigned long long)__double_as_longlong(lo + kFixL) - (unsigned long long)__double_as_longlong(kFixL);
                }
#endif
            } else
#endif
#if STX_K_STATS_TILE
            {
                // (partial tile: the same two accumulators per parity of i as the full-tile path, rows past the end add nothing)
                double a1[2] = {0.0, 0.0}, a2[2] = {0.0, 0.0};
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                    const int row = srow + 3 * i;
                    if (row < rows) {
                        const float v = outstage[row * kOutRow + sbin];
                        if (row < keep) dst[i * kGStat] = v;
                        const double vd = (double)v;
                        a1[i & 1] += vd;
                        a2[i & 1] = fma(vd, vd, a2[i & 1]);
                    }
                }
                const double t1 = a1[0] + a1[1], t2 = a2[0] + a2[1];
                const double hi = t2 + kFixH, lo = t2 - (hi - kFixH);
                s1 += (unsigned long long)__double_as_longlong(t1 + kFix1) - (unsigned long long)__double_as_longlong(kFix1);
                s2h += (unsigned long long)__double_as_longlong(hi) - (unsigned long long)__double_as_longlong(kFixH);
                s2l += (unsigned long long)__double_as_longlong(lo + kFixL) - (unsigned long long)__double_as_longlong(kFixL);
            }
#else
#pragma unroll
            for (int i = 0; i < 11; ++i) {
                const int row = srow + 3 * i;
                if (row < rows) {
                    const float v = outstage[row * kOutRow + sbin];
                    if (row < keep) dst[i * kGStat] = v;
                    const double vd = (double)v;
                    const double sq = vd * vd, hi = sq + kFixH, lo = sq - (hi - kFixH);      // all exact
                    s1 += (unsigned long long)__double_as_longlong(vd + kFix1) - (unsigned long long)__double_as_longlong(kFix1);
                    s2h += (unsigned long long)__double_as_longlong(hi) - (unsigned long long)__double_as_longlong(kFixH);
                    s2l += (unsigned long long)__double_as_longlong(lo + kFixL) - (unsigned long long)__double_as_longlong(kFixL);
                }
            }
#endif
        }
        if (cur.last) {
            // ---- last tile of the item: the group's 3 row groups -> partials[item] (ordered, integer: exact) ----
            if (tl < kGStat) { red[tl] = s1; red[kGStat + tl] = s2h; red[2 * kGStat + tl] = s2l; }
            group_bar(g);
            if (tl < kStatWords) {
                const int which = tl / kMel, m = tl - which * kMel;
                cur.part[tl] = (long long)(red[which * kGStat + m] + red[which * kGStat + kMel + m] + red[which * kGStat + 2 * kMel + m]);
            }
            s1 = 0; s2h = 0; s2l = 0;
#if STX_K_CLIPSTATS
            if (cstat) {
                // ---- count the item; the group that completes the clip reduces its statistics (integer sums of the items'
                // partials: the same bits whichever group does it) and leaves mean / 1/std per bin for k_normalize ----
                __threadfence();                     // this item's partials are visible device-wide before it is counted
                group_bar(g);
                if (tl == 0) {
                    const unsigned old = atomicAdd(done + cur.b, 1u);
                    const int last = (old + 1u == (unsigned)cur.nitems);
                    if (last) done[cur.b] = 0u;      // self-cleaning: the counters are zero again when the kernel ends
                    sg.is_last = last;
                }
                group_bar(g);
                if (sg.is_last) {
                    __threadfence();
                    if (tl < kStatWords) {
                        const long long* p = partials + (size_t)cur.b * chunks_per_clip * kStatWords + tl;
                        const int nit = cur.nitems;
                        long long acc = 0;
                        int c = 0;
                        for (; c + 8 <= nit; c += 8) {                       // 8 independent L2 loads in flight
                            long long v[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) v[u] = __ldcg(p + (size_t)(c + u) * kStatWords);
#pragma unroll
                            for (int u = 0; u < 8; ++u) acc += v[u];
                        }
                        for (; c < nit; ++c) acc += __ldcg(p + (size_t)c * kStatWords);
                        red[tl] = (unsigned long long)acc;
                    }
                    group_bar(g);
                    if (tl < kMel) {
                        const int T_all = min(1 + (cur.n - kFrame) / kHop, chunks_per_clip * chunk_frames);
                        float mh, ml, rs;
                        clip_stats((long long)red[tl], (long long)red[kMel + tl], (long long)red[2 * kMel + tl], T_all, mh, ml, rs);
                        float* cs = cstat + (size_t)cur.b * kStatWords;
                        cs[tl] = mh; cs[kMel + tl] = ml; cs[2 * kMel + tl] = rs;
                    }
                }
            }
#endif
        }
        slot ^= 1;
      }
        if (sg.desc[slot].valid) convert(sg.desc[slot]);
        KTRACE_PRE(6);
        group_bar(g);                               // d ready; landing zone and staged rows consumed
        KTRACE(6);
        first_trip = false;
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
#if STX_K_TRACE == 1
    if (tl < 8) atomicAdd(&g_ktrace[tl], (unsigned long long)sg.tacc[tl]);
#elif STX_K_TRACE == 2
    if (tl < 64) atomicAdd(&g_ktrace[tl], (unsigned long long)sg.tw[tl >> 3][tl & 7]);
#endif
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem_base), "r"(kTmemCols) : "memory");
    }
}

// The work items of k_frames_duo, compacted: sched[0] = number of non-empty items, sched[1 + i] = b * chunks_per_clip + chunk in
// clip-major order.  One CTA; a block-wide exclusive scan of the per-clip chunk counts, 1024 clips at a time.  (Lengths live
// on the device, so the host cannot do this; without it the static round-robin over ALL items leaves the busiest group of a
// ragged batch with a third more tiles than the average.)
__global__ void __launch_bounds__(1024)
k_schedule(const int* __restrict__ lengths, int B, int chunk_frames, int chunks_per_clip, int* __restrict__ sched) {
    // programmatic dependent launch: k_frames_duo may start its prologue (TMEM allocation, tables) right away; it waits for
    // this grid's completion (griddepcontrol.wait) before it reads the schedule
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ int s_warp[32];
    __shared__ int s_off[1024], s_cnt[1024];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int b = base + tid;
        int cnt = 0;
        if (b < B) {
            const int n = lengths[b];
            const int T = n >= kFrame ? 1 + (n - kFrame) / kHop : 0;
            cnt = (T + chunk_frames - 1) / chunk_frames;
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
            s_warp[lane] = w;                        // inclusive scan of the warp totals
        }
        __syncthreads();
        s_off[tid] = s_carry + (warp ? s_warp[warp - 1] : 0) + incl - cnt;
        s_cnt[tid] = cnt;
        __syncthreads();
        // one warp per clip writes that clip's items: coalesced, and as parallel as the batch is wide
        const int nclips = min(1024, B - base);
        for (int i = warp; i < nclips; i += 32)
            for (int c = lane; c < s_cnt[i]; c += 32) sched[1 + s_off[i] + c] = (base + i) * chunks_per_clip + c;
        __syncthreads();
        if (tid == 0) s_carry += s_warp[31];
        __syncthreads();
    }
    if (tid == 0) sched[0] = s_carry;
}

// in-place CMVN + padding rows + mask.  One thread per float4 of a clip's [T_pad, 80] block.
//   rows t < T            normalised features
//   rows T <= t < T2      padding_value (T2 = T rounded up to even: the half of the last stacked frame of an odd clip)
//   rows t >= T2          tail_value (= padding_value for the extractor, …seamless_m4t.py:267-275; 0 for the trainer's
//                         collate, R/training/trainer_unfreeze.py:902)
//   mask_mode 0: int32, mask[j] = (2 j + 1 < T)  (…seamless_m4t.py:292-293);  1: int64, mask[j] = (2 j < T), i.e. every
//   stacked frame the per-clip extractor call returned (R/training/trainer_unfreeze.py:904-908)
__global__ void __launch_bounds__(256)
k_normalize(const int* __restrict__ lengths, const long long* __restrict__ partials, int chunk_frames, int chunks_per_clip,
            int T_pad, float padding_value, float tail_value, int normalize, float* __restrict__ out,
            void* __restrict__ mask, int mask_mode, const float* __restrict__ cstat) {
    __shared__ __align__(16) float s_mean_hi[kMel], s_mean_lo[kMel], s_rstd_f[kMel];
    __shared__ long long s_sum[kStatWords];
    const int b = blockIdx.y;
    const int n = lengths[b];
#if STX_K_NORM_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");      // k_frames_duo (the programmatic primary) has completed
#endif
    // the statistics are over ALL frames of the clip (all that were processed: a clip longer than the max_length the call was
    // sized for is cut there, in every kernel alike)
    const int T_all = min(n >= kFrame ? 1 + (n - kFrame) / kHop : 0, chunks_per_clip * chunk_frames);
    const int T = min(T_all, T_pad);
    const int T2 = min((T + 1) & ~1, T_pad);
    // mean and 1/sqrt(var + 1e-7) per bin, var with ddof = 1 (…seamless_m4t.py:257-262): integer sums of the chunk
    // partials are exact and order-independent, so every CTA of the clip (and every batch) gets the same bits
    if (normalize && T_all > 0 && cstat) {
        // the statistics were reduced inside k_frames_duo by the group that completed the clip
        if (threadIdx.x < kMel) {
            const float* cs = cstat + (size_t)b * kStatWords;
            s_mean_hi[threadIdx.x] = cs[threadIdx.x];
            s_mean_lo[threadIdx.x] = cs[kMel + threadIdx.x];
            s_rstd_f[threadIdx.x] = cs[2 * kMel + threadIdx.x];
        }
    } else if (normalize && T_all > 0) {
        const int nchunks = min((T_all + chunk_frames - 1) / chunk_frames, chunks_per_clip);
        if (threadIdx.x < kStatWords) {
            long long acc = 0;
            const long long* p = partials + (size_t)b * chunks_per_clip * kStatWords + threadIdx.x;
            int c = 0;
            for (; c + 8 <= nchunks; c += 8) {                            // 8 independent L2 loads in flight
                long long v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldg(p + (size_t)(c + u) * kStatWords);
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += v[u];
            }
            for (; c < nchunks; ++c) acc += __ldg(p + (size_t)c * kStatWords);
            s_sum[threadIdx.x] = acc;
        }
        __syncthreads();
        if (threadIdx.x < kMel) {
            const int m = threadIdx.x;
            clip_stats(s_sum[m], s_sum[kMel + m], s_sum[2 * kMel + m], T_all, s_mean_hi[m], s_mean_lo[m], s_rstd_f[m]);
        }
    }
    __syncthreads();
    const int quads = T_pad * (kMel / 4);
    float4* o4 = reinterpret_cast<float4*>(out + (size_t)b * T_pad * kMel);
#if STX_K_NORM_COLS
    // 240 of the 256 threads: thread i owns the column quad i % 20 of the rows blockIdx.x * 12 + i / 20 + k * gridDim.x * 12, so
    // its mean / 1/std words stay in registers and the loop has neither shared-memory reads nor a division; consecutive threads
    // still touch consecutive float4s (a row is 20 of them)
    if (threadIdx.x < 240) {
        const int m = (threadIdx.x % (kMel / 4)) * 4;
        float4 mh = make_float4(0.f, 0.f, 0.f, 0.f), ml = mh, rs = mh;
        if (normalize && T > 0) {
            mh = *reinterpret_cast<const float4*>(&s_mean_hi[m]);
            ml = *reinterpret_cast<const float4*>(&s_mean_lo[m]);
            rs = *reinterpret_cast<const float4*>(&s_rstd_f[m]);
        }
        const int row_step = gridDim.x * 12;
        int t = blockIdx.x * 12 + threadIdx.x / (kMel / 4);
        for (int q = blockIdx.x * 240 + threadIdx.x; q < quads; q += gridDim.x * 240, t += row_step) {
            float4 v;
            if (t < T) {
                if (!normalize) continue;
                v = o4[q];
                // float32 with a two-word mean: (x - mean_hi) is exact or nearly so (same binade), mean_lo restores the bits the
                // float32 mean lost, and the float32 1/std costs 6e-8 of a result of magnitude <= 10: |error| < 1e-6
                v.x = ((v.x - mh.x) - ml.x) * rs.x;
                v.y = ((v.y - mh.y) - ml.y) * rs.y;
                v.z = ((v.z - mh.z) - ml.z) * rs.z;
                v.w = ((v.w - mh.w) - ml.w) * rs.w;
            } else {
                const float f = t < T2 ? padding_value : tail_value;
                v = make_float4(f, f, f, f);
            }
            o4[q] = v;
        }
    }
#else
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
        const int t = q / (kMel / 4), m = (q - t * (kMel / 4)) * 4;
        float4 v;
        if (t < T) {
            if (!normalize) continue;
            v = o4[q];
            // float32 with a two-word mean: (x - mean_hi) is exact or nearly so (same binade), mean_lo restores the bits the
            // float32 mean lost, and the float32 1/std costs 6e-8 of a result of magnitude <= 10: |error| < 1e-6, with no
            // quarter-rate F2F conversions (8 per float4 in the float64 form: 31.8 -> 24.9 us on cfg2)
            const float4 mh = *reinterpret_cast<const float4*>(&s_mean_hi[m]);
            const float4 ml = *reinterpret_cast<const float4*>(&s_mean_lo[m]);
            const float4 rs = *reinterpret_cast<const float4*>(&s_rstd_f[m]);
            v.x = ((v.x - mh.x) - ml.x) * rs.x;
            v.y = ((v.y - mh.y) - ml.y) * rs.y;
            v.z = ((v.z - mh.z) - ml.z) * rs.z;
            v.w = ((v.w - mh.w) - ml.w) * rs.w;
        } else {
            const float f = t < T2 ? padding_value : tail_value;
            v = make_float4(f, f, f, f);
        }
        o4[q] = v;
    }
#endif
    if (mask) {
        const int rows = T_pad / 2;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < rows; j += gridDim.x * blockDim.x) {
            if (mask_mode == 0) static_cast<int*>(mask)[(size_t)b * rows + j] = (2 * j + 1 < T) ? 1 : 0;
            else static_cast<long long*>(mask)[(size_t)b * rows + j] = (2 * j < T) ? 1 : 0;
        }
    }
}

// CMVN + padding + LayerNorm(160) + TF32 hi / lo split in ONE pass over the raw log-mel: the encoder's input stage
// (Wav2Vec2BertFeatureProjection, TF/models/wav2vec2_bert/modeling_wav2vec2_bert.py:118-130) takes LayerNorm(input_features),
// so when the caller wants the projected hidden states the normalised features need never be written and re-read
// (SURVEY.md 8f row 2).  One warp per stacked row (frames 2 j and 2 j + 1 = 160 values, 5 per lane); statistics prologue and
// CMVN arithmetic exactly as in k_normalize (the optional `feat` output is bit-identical to stx_fbank_k's); raw and feat may
// be the same buffer.
__global__ void __launch_bounds__(256)
k_norm_ln_split(const int* __restrict__ lengths, const long long* __restrict__ partials, int chunk_frames, int chunks_per_clip,
                int T_pad, float padding_value, const float* raw, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, float* __restrict__ planes, size_t plane_stride,
                float* feat, int* __restrict__ mask, const float* __restrict__ cstat) {
    __shared__ __align__(16) float s_mean_hi[kMel], s_mean_lo[kMel], s_rstd_f[kMel];
    __shared__ long long s_sum[kStatWords];
    const int b = blockIdx.y;
    const int n = lengths[b];
    const int T_all = min(n >= kFrame ? 1 + (n - kFrame) / kHop : 0, chunks_per_clip * chunk_frames);
    const int T = min(T_all, T_pad);
    if (T_all > 0 && cstat) {
        if (threadIdx.x < kMel) {
            const float* cs = cstat + (size_t)b * kStatWords;
            s_mean_hi[threadIdx.x] = cs[threadIdx.x];
            s_mean_lo[threadIdx.x] = cs[kMel + threadIdx.x];
            s_rstd_f[threadIdx.x] = cs[2 * kMel + threadIdx.x];
        }
    } else if (T_all > 0) {
        const int nchunks = (T_all + chunk_frames - 1) / chunk_frames;
        if (threadIdx.x < kStatWords) {
            long long acc = 0;
            const long long* p = partials + (size_t)b * chunks_per_clip * kStatWords + threadIdx.x;
            for (int c = 0; c < nchunks; ++c) acc += __ldg(p + (size_t)c * kStatWords);
            s_sum[threadIdx.x] = acc;
        }
        __syncthreads();
        if (threadIdx.x < kMel) {
            const int m = threadIdx.x;
            clip_stats(s_sum[m], s_sum[kMel + m], s_sum[2 * kMel + m], T_all, s_mean_hi[m], s_mean_lo[m], s_rstd_f[m]);
        }
    }
    __syncthreads();
    const int rows = T_pad / 2;
    const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < rows; j += warps) {
        const size_t row = (size_t)b * rows + j;
        const float* src = raw + row * (2 * kMel);
        float f[5];
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int i = lane + 32 * k;
            const int t = 2 * j + (i >= kMel), m = i - (i >= kMel ? kMel : 0);
            f[k] = t < T ? ((src[i] - s_mean_hi[m]) - s_mean_lo[m]) * s_rstd_f[m] : padding_value;
            if (feat) feat[row * (2 * kMel) + i] = f[k];
            sum += f[k];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mu = sum * (1.0f / (2 * kMel));
        float q = 0.0f;
#pragma unroll
        for (int k = 0; k < 5; ++k) { const float d = f[k] - mu; q = fmaf(d, d, q); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.0f / (2 * kMel)) + eps);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int i = lane + 32 * k;
            const float v = (f[k] - mu) * rstd * __ldg(gamma + i) + __ldg(beta + i);
            const float h = __int_as_float(__float_as_int(v) & 0xffffe000);
            planes[row * (2 * kMel) + i] = h;
            planes[plane_stride + row * (2 * kMel) + i] = v - h;
        }
        if (mask && lane == 0) mask[row] = (2 * j + 1 < T) ? 1 : 0;
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
        static KTables h;
        const std::vector<double>& w = k_window();
        // constant-bank tables; FFT512 of the scaled window by direct summation in long double
        static double win[16][25];
        static double2 tw[16][16], wh[16][16];
#if STX_K_POW_PRESCALE && !STX_K_P1_F32 && !STX_K_P2_F32
        const double kWinScale = std::ldexp(32768.0, -448);     // 2^15 (Kaldi's int16 scale) * 2^-448 (see STX_K_POW_PRESCALE)
#else
        const double kWinScale = 32768.0;
#endif
        for (int n2 = 0; n2 < 16; ++n2)
            for (int n1 = 0; n1 < 25; ++n1) {
                win[n2][n1] = w[16 * n1 + n2] * kWinScale;
            }
        for (int n2 = 0; n2 < 16; ++n2)
            for (int k1 = 0; k1 < 16; ++k1) {
                const double ang = -2.0 * M_PI * double(n2 * k1) / 512.0;
                tw[n2][k1] = make_double2(std::cos(ang), std::sin(ang));
            }
        auto what = [&](int k) {
            long double re = 0.0L, im = 0.0L;
            for (int i = 0; i < kFrame; ++i) {
                const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)((i * k) % 512) / 512.0L;
                re += (long double)(w[i] * 32768.0) * cosl(ang);
                im += (long double)(w[i] * 32768.0) * sinl(ang);
            }
            return make_double2((double)re * (kWinScale / 32768.0), (double)im * (kWinScale / 32768.0));
        };
        for (int k1 = 0; k1 < 16; ++k1)
            for (int k2 = 0; k2 < 16; ++k2) wh[k1][k2] = what(k1 + 32 * k2);
        for (int k2 = 0; k2 < 8; ++k2) wh[0][8 + k2] = what(16 + 32 * k2);     // row 0: bins 32 k2 and 16 + 32 k2
#if STX_K_P1_F32 || STX_K_P2_F32
        {
            static float win_f[16][25];
            static float2 tw_f[16][16], wh_f[16][16];
            for (int i = 0; i < 16; ++i) {
                for (int j = 0; j < 25; ++j) win_f[i][j] = (float)win[i][j];
                for (int j = 0; j < 16; ++j) {
                    tw_f[i][j] = make_float2((float)tw[i][j].x, (float)tw[i][j].y);
                    wh_f[i][j] = make_float2((float)wh[i][j].x, (float)wh[i][j].y);
                }
            }
            STX_CUDA(cudaMemcpyToSymbol(c_win_f, win_f, sizeof(win_f)));
            STX_CUDA(cudaMemcpyToSymbol(c_tw_f, tw_f, sizeof(tw_f)));
            STX_CUDA(cudaMemcpyToSymbol(c_wh_f, wh_f, sizeof(wh_f)));
        }
#endif
        STX_CUDA(cudaMemcpyToSymbol(c_win, win, sizeof(win)));
        STX_CUDA(cudaMemcpyToSymbol(c_tw, tw, sizeof(tw)));
        STX_CUDA(cudaMemcpyToSymbol(c_wh, wh, sizeof(wh)));
        // mel filters, padded per slot to mel_len(slot) weights ending at or before bin 255
        const std::vector<double>& fb = k_mel();
        for (int m = 0; m < kMel; ++m) {
            const int slot = m / 16, wrp = m % 16, L = mel_len(slot);
            int lo = -1, hi = -1;
            for (int k = 0; k <= STX_K_NFFT / 2; ++k)
                if (fb[size_t(k) * kMel + m] != 0.0) { if (lo < 0) lo = k; hi = k; }
            if (lo < 0 || hi - lo + 1 > L || hi > 255) { set_error("mel filter %d does not fit its slot", m); return STX_EINVAL; }
            if (lo + L > 256) lo = 256 - L;
            h.melfirst[m] = lo;
            for (int q = 0; q < L; ++q) h.melw[mel_off(slot) + wrp * L + q] = float(fb[size_t(lo + q) * kMel + m]);
        }
        // the same filters in half slots of 8 (STX_K_MEL_EXACT): mel_len10 bins read, weights padded to a multiple of four
        for (int i = 0; i < kMelWeights; ++i) h.melw10[i] = 0.0f;
        for (int m = 0; m < kMel; ++m) {
            const int j = m / 8, w8 = m % 8, L = mel_len10(j), PAD = mel_pad10(j);
            int lo = -1, hi = -1;
            for (int k = 0; k <= STX_K_NFFT / 2; ++k)
                if (fb[size_t(k) * kMel + m] != 0.0) { if (lo < 0) lo = k; hi = k; }
            if (lo < 0 || hi - lo + 1 > L || hi > 255) { set_error("mel filter %d does not fit its half slot", m); return STX_EINVAL; }
            if (lo + L > 256) lo = 256 - L;
            h.melfirst10[m] = lo;
            for (int q = 0; q < L; ++q) h.melw10[mel_off10(j) + w8 * PAD + q] = float(fb[size_t(lo + q) * kMel + m]);
        }
        // the generated mel stage (mel_k.cuh) carries its weights as immediates: they must be this library's table
        {
            int at = 0;
            for (int m = 0; m < kMel; ++m) {
                int lo = -1, hi = -1;
                for (int k = 0; k <= STX_K_NFFT / 2; ++k)
                    if (fb[size_t(k) * kMel + m] != 0.0) { if (lo < 0) lo = k; hi = k; }
                if (lo != melk::kMelFirstBin[m] || hi != melk::kMelLastBin[m]) { set_error("mel_k.cuh: filter %d spans bins %d..%d, the table says %d..%d (rerun tools/gen_mel_k.py)", m, melk::kMelFirstBin[m], melk::kMelLastBin[m], lo, hi); return STX_EINVAL; }
                for (int k = lo; k <= hi; ++k, ++at)
                    if (melk::kMelWeightsFlat[at] != float(fb[size_t(k) * kMel + m])) { set_error("mel_k.cuh: weight of filter %d at bin %d differs from the table (rerun tools/gen_mel_k.py)", m, k); return STX_EINVAL; }
            }
        }
        STX_CUDA(cudaMemcpyToSymbol(c_melw, h.melw, sizeof(h.melw)));
        STX_CUDA(cudaMemcpyToSymbol(c_melfirst, h.melfirst, sizeof(h.melfirst)));
        KTables* d = nullptr;
        STX_CUDA(cudaMalloc(&d, sizeof(KTables)));
        STX_CUDA(cudaMemcpy(d, &h, sizeof(KTables), cudaMemcpyHostToDevice));
        STX_CUDA(cudaFuncSetAttribute(k_frames<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        STX_CUDA(cudaFuncSetAttribute(k_frames<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        STX_CUDA(cudaFuncSetAttribute(k_frames_duo<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemDuo)));
        STX_CUDA(cudaFuncSetAttribute(k_frames_duo<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemDuo)));
        g_tab[dev] = d;
    }
    *out = g_tab[dev];
    return 0;
}

inline int frames_of(int n) { return n >= kFrame ? 1 + (n - kFrame) / kHop : 0; }

// Frames per CTA (a multiple of the 32-frame tile): with one CTA per SM the grid runs in waves of `sms` CTAs, so
// pick the chunk that minimises waves * tiles-per-chunk for B clips of max_frames; ties go to the larger chunk
// (fewer prologues).  Lengths live on the device, so ragged batches are sized by their longest clip.
inline int pick_chunk(int B, int max_frames, int sms) {
    int best = kMinChunk;
    long long best_cost = -1;
    for (int chunk = kMinChunk; chunk <= 512; chunk += kTile) {
        const long long ctas = (long long)B * ((max_frames + chunk - 1) / chunk);
        const long long cost = ((ctas + sms - 1) / sms) * (chunk / kTile);
        if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = chunk; }
    }
    return best;
}
// k_frames_duo is persistent: 2 * sms groups stream through B * ceil(max_frames / chunk) items, so the chunk (>= kMinChunk
// frames: the partials workspace is sized for that) minimises items-per-group * tiles-per-item; ties go to the larger chunk
inline int pick_chunk_duo(int B, int max_frames, int sms) {
    int best = kMinChunk;
    long long best_cost = -1;
    for (int chunk = kMinChunk; chunk <= 512; chunk += kTile) {
        const long long items = (long long)B * ((max_frames + chunk - 1) / chunk);
        const long long groups = 2 * std::max<long long>(1, std::min<long long>(sms, (items + 1) / 2));   // items = 0: clips shorter than a frame
        const long long cost = ((items + groups - 1) / groups) * (chunk / kTile);
        if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = chunk; }
    }
    return best;
}
// k_frames_duo is launched as the programmatic dependent of k_schedule: its CTAs start while the (one-CTA) scheduler still
// runs and block in griddepcontrol.wait only where they first read the schedule
template <typename K, typename... Args>
int launch_dependent(const char* name, K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const bool prof = g_profile.load(std::memory_order_relaxed) != 0;
    if (prof) profile_before(name, st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    if (prof) profile_after(st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return cuda_fail(e, name);
    return 0;
}
// Per (device, stream) completion counters of k_frames_duo (one per clip).  They are zero whenever no call is in flight on
// the stream: allocated and zeroed once, and the kernel resets every counter it has driven to its final value (calls on one
// stream are serialised, so one set per stream is enough; the memory is never freed, 256 KB per stream that ever ran recipe K).
constexpr int kMaxClips = 65536;
std::mutex g_state_mutex;
std::vector<std::pair<std::pair<int, cudaStream_t>, unsigned*>> g_state;
int get_counters(cudaStream_t st, unsigned** out) {
    int dev = 0;
    STX_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_state_mutex);
    for (auto& e : g_state)
        if (e.first.first == dev && e.first.second == st) { *out = e.second; return 0; }
    unsigned* p = nullptr;
    STX_CUDA(cudaMalloc(&p, sizeof(unsigned) * kMaxClips));
    STX_CUDA(cudaMemsetAsync(p, 0, sizeof(unsigned) * kMaxClips, st));
    g_state.push_back({{dev, st}, p});
    *out = p;
    return 0;
}
inline bool use_duo() {
    static const bool v = [] { const char* e = std::getenv("STX_K_SINGLE"); return !(e && e[0] == '1'); }();
    return v;
}
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace
}  // namespace stx

extern "C" {

#if STX_K_TRACE
// debug builds only (tools/k_phase_trace.py): cycles per phase summed over all groups, then cleared
int stx_debug_ktrace(unsigned long long* host16) {       // (64 words)
    using namespace stx;
    STX_CUDA(cudaDeviceSynchronize());
    STX_CUDA(cudaMemcpyFromSymbol(host16, g_ktrace, sizeof(unsigned long long) * 64));
    static const unsigned long long zero[64] = {0};
    STX_CUDA(cudaMemcpyToSymbol(g_ktrace, zero, sizeof(zero)));
    return 0;
}
#endif

int stx_fbank_k_workspace(int B, int max_length, size_t* bytes) {
    using namespace stx;
    if (B < 0 || max_length == INT32_MIN || !bytes) { set_error("stx_fbank_k_workspace: bad argument"); return STX_EINVAL; }
    if (max_length < 0) max_length = -max_length;          // the "uniform batch" form of stx_fbank_k
    const int chunks = (frames_of(max_length) + kMinChunk - 1) / kMinChunk;
    // per-chunk statistics partials, then the schedule of k_frames_duo (item count + one int per item), then the per-clip
    // statistics (mean hi / lo, 1/std per bin)
    *bytes = align256(size_t(B) * std::max(chunks, 1) * kStatWords * sizeof(long long)) +
             align256(sizeof(int) * (1 + size_t(B) * std::max(chunks, 1))) +
             align256(size_t(B) * kStatWords * sizeof(float));
    return 0;
}

// What the second pass of a fused consumer needs to know about the first (stx_fbank_k_projection)
struct FrontInfo { int chunk_frames, chunks, sms; const long long* partials; const float* cstat; };

static int fbank_k_impl(const float* d_pcm, const int64_t* d_offsets, const int32_t* d_lengths, int B, int max_length,
                        const float* d_peak, int T_pad, float padding_value, float tail_value, int normalize,
                        float* d_out, void* d_mask, int mask_mode, void* d_ws, size_t ws_bytes, void* stream,
                        FrontInfo* info = nullptr) {
    using namespace stx;
    // max_length < 0: the caller promises that EVERY clip has exactly -max_length samples
    const bool uniform = max_length < 0;
    if (max_length == INT32_MIN) { set_error("stx_fbank_k: bad max_length"); return STX_EINVAL; }
    if (uniform) max_length = -max_length;
    if (B < 0 || T_pad < 0 || (T_pad & 1)) { set_error("stx_fbank_k: B >= 0 and even T_pad >= 0 required"); return STX_EINVAL; }
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

    const int max_frames = frames_of(max_length);
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        STX_CUDA(cudaGetDevice(&dev));
        STX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const bool duo = use_duo();
    int chunk_frames = duo ? pick_chunk_duo(B, max_frames, sms) : pick_chunk(B, max_frames, sms);
    if (const char* e = std::getenv("STX_K_CHUNK")) {        // development: force the frames per work item (a multiple of 32, >= 64)
        const int v = std::atoi(e);
        if (v >= kMinChunk && v % kTile == 0) chunk_frames = v;
    }
    const int chunks = std::max((max_frames + chunk_frames - 1) / chunk_frames, 1);
    long long* partials = static_cast<long long*>(d_ws);
    const int chunks64 = std::max((frames_of(max_length) + kMinChunk - 1) / kMinChunk, 1);
    int* sched = reinterpret_cast<int*>(static_cast<unsigned char*>(d_ws) + align256(size_t(B) * chunks64 * kStatWords * sizeof(long long)));
    float* cstat = nullptr;
    unsigned* done = nullptr;
#if STX_K_CLIPSTATS
    if (duo && normalize && frames_of(max_length) > 0) {
        cstat = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sched) + align256(sizeof(int) * (1 + size_t(B) * chunks64)));
        if (int rc = get_counters(st, &done)) return rc;
    }
#endif
    if (frames_of(max_length) > 0) {
        const int duo_grid = (int)std::min<long long>(sms, ((long long)B * chunks + 1) / 2);
        // uniform batch (the caller's promise): no empty items, the plain round-robin is balanced, and the 3.6 us of
        // k_schedule are saved; a wrong promise only costs load balance, never correctness
        const bool scheduled = duo && !uniform && B > 1;
        if (scheduled) STX_LAUNCH(k_schedule, dim3(1), dim3(1024), 0, st, d_lengths, B, chunk_frames, chunks, sched);
        const int* sched_arg = scheduled ? sched : nullptr;
        if (duo) {
            const long long* off = reinterpret_cast<const long long*>(d_offsets);
            const int rc = d_peak
                ? launch_dependent("k_frames_duo<true>", k_frames_duo<true>, dim3(duo_grid), dim3(kThreads), sizeof(SmemDuo), st,
                                   d_pcm, off, d_lengths, d_peak, tab, B, T_pad, chunk_frames, chunks, d_out, partials, sched_arg, done, cstat)
                : launch_dependent("k_frames_duo<false>", k_frames_duo<false>, dim3(duo_grid), dim3(kThreads), sizeof(SmemDuo), st,
                                   d_pcm, off, d_lengths, d_peak, tab, B, T_pad, chunk_frames, chunks, d_out, partials, sched_arg, done, cstat);
            if (rc) return rc;
        } else if (d_peak) {
            STX_LAUNCH(k_frames<true>, dim3(chunks, B), dim3(kThreads), sizeof(Smem), st,
                       d_pcm, reinterpret_cast<const long long*>(d_offsets), d_lengths, d_peak, tab, T_pad,
                       chunk_frames, chunks, d_out, partials);
        } else {
            STX_LAUNCH(k_frames<false>, dim3(chunks, B), dim3(kThreads), sizeof(Smem), st,
                       d_pcm, reinterpret_cast<const long long*>(d_offsets), d_lengths, d_peak, tab, T_pad,
                       chunk_frames, chunks, d_out, partials);
        }
    }
    if (info) {
        info->chunk_frames = chunk_frames; info->chunks = chunks; info->sms = sms; info->partials = partials; info->cstat = cstat;
        return 0;
    }
    const int quads = T_pad * (kMel / 4);
    // every CTA pays a ~2 us prologue (partials -> mean, 1/std), so the grid is ONE wave of 8 CTAs per SM, not more
    // (cfg2: 64 x 16 CTAs 20.9 us, 64 x 64 CTAs 24.9 us), and never more CTAs than 256-thread groups of float4s
    int gx = std::max(1, std::min((quads + 255) / 256, std::max(1, (8 * sms) / B)));
    if (const char* e = std::getenv("STX_KN_GX")) { const int v = std::atoi(e); if (v > 0) gx = v; }     // development: CTAs per clip
#if STX_K_NORM_PDL
    if (duo)
        return launch_dependent("k_normalize", k_normalize, dim3(gx, B), dim3(256), 0, st, d_lengths, (const long long*)partials,
                                chunk_frames, chunks, T_pad, padding_value, tail_value, normalize, d_out, d_mask, mask_mode,
                                (const float*)cstat);
#endif
    STX_LAUNCH(k_normalize, dim3(gx, B), dim3(256), 0, st, d_lengths, partials, chunk_frames, chunks, T_pad, padding_value,
               tail_value, normalize, d_out, d_mask, mask_mode, (const float*)cstat);
    return 0;
}

int stx_fbank_k(const float* d_pcm, const int64_t* d_offsets, const int32_t* d_lengths, int B, int max_length,
                const float* d_peak, int T_pad, float padding_value, int normalize, float* d_out,
                int32_t* d_mask, void* d_ws, size_t ws_bytes, void* stream) {
    return fbank_k_impl(d_pcm, d_offsets, d_lengths, B, max_length, d_peak, T_pad, padding_value, padding_value, normalize,
                        d_out, d_mask, 0, d_ws, ws_bytes, stream);
}

int stx_fbank_k_collate(const float* d_pcm, const int64_t* d_offsets, const int32_t* d_lengths, int B, int max_length,
                        int T_pad, float padding_value, float* d_out, int64_t* d_mask, void* d_ws, size_t ws_bytes,
                        void* stream) {
    return fbank_k_impl(d_pcm, d_offsets, d_lengths, B, max_length, nullptr, T_pad, padding_value, 0.0f, 1,
                        d_out, d_mask, 1, d_ws, ws_bytes, stream);
}

int stx_fbank_k_projection_workspace(int B, int max_length, int T_pad, int out_dim, int want_features, size_t* bytes) {
    using namespace stx;
    if (B < 0 || T_pad < 0 || (T_pad & 1) || out_dim <= 0 || !bytes) { set_error("stx_fbank_k_projection_workspace: bad argument"); return STX_EINVAL; }
    size_t front = 0;
    if (int rc = stx_fbank_k_workspace(B, max_length, &front)) return rc;
    const size_t rows = size_t(B) * (T_pad / 2);
    *bytes = front + (want_features ? 0 : align256(rows * 2 * kMel * sizeof(float))) +
             2 * align256(rows * 2 * kMel * sizeof(float)) + 2 * align256(size_t(out_dim) * 2 * kMel * sizeof(float));
    return 0;
}

int stx_fbank_k_projection(const float* d_pcm, const int64_t* d_offsets, const int32_t* d_lengths, int B, int max_length,
                           const float* d_peak, int T_pad, float padding_value, const float* d_ln_weight,
                           const float* d_ln_bias, float eps, const float* d_weight, const float* d_bias, int out_dim,
                           float* d_hidden, float* d_features, int32_t* d_mask, void* d_ws, size_t ws_bytes, void* stream) {
    using namespace stx;
    if (B < 0 || T_pad < 0 || (T_pad & 1) || out_dim <= 0) { set_error("stx_fbank_k_projection: bad argument"); return STX_EINVAL; }
    if (B == 0 || T_pad == 0) return 0;
    if (!d_ln_weight || !d_ln_bias || !d_weight || !d_hidden || !d_ws) { set_error("stx_fbank_k_projection: null pointer"); return STX_EINVAL; }
    size_t need = 0, front = 0;
    if (int rc = stx_fbank_k_projection_workspace(B, max_length, T_pad, out_dim, d_features != nullptr, &need)) return rc;
    if (ws_bytes < need) { set_error("stx_fbank_k_projection: workspace %zu < %zu bytes", ws_bytes, need); return STX_ENOSPACE; }
    stx_fbank_k_workspace(B, max_length, &front);
    const size_t rows = size_t(B) * (T_pad / 2);
    unsigned char* p = static_cast<unsigned char*>(d_ws) + front;
    float* raw = d_features;
    if (!raw) { raw = reinterpret_cast<float*>(p); p += align256(rows * 2 * kMel * sizeof(float)); }
    float* a_planes = reinterpret_cast<float*>(p);  p += 2 * align256(rows * 2 * kMel * sizeof(float));
    float* b_planes = reinterpret_cast<float*>(p);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FrontInfo info = {};
    if (int rc = fbank_k_impl(d_pcm, d_offsets, d_lengths, B, max_length, d_peak, T_pad, padding_value, padding_value, 1, raw,
                              nullptr, 0, d_ws, front, stream, &info)) return rc;
    const int gx = std::max(1, std::min((T_pad / 2 + 7) / 8, std::max(1, (8 * info.sms) / B)));
    STX_LAUNCH(k_norm_ln_split, dim3(gx, B), dim3(256), 0, st, d_lengths, info.partials, info.chunk_frames, info.chunks, T_pad,
               padding_value, (const float*)raw, d_ln_weight, d_ln_bias, eps, a_planes, rows * 2 * kMel, d_features, d_mask,
               info.cstat);
    return project_from_planes(a_planes, (int)rows, 2 * kMel, d_weight, d_bias, out_dim, b_planes, d_hidden, st);
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
