// Cosine scoring of L2-normalised embeddings for sm_100a.
//
// Replaces AudioTextProcessor.compute_similarity (R/processor.py:148-159) and the
// F.normalize(p=2, dim=1) + (a * b).sum(dim=1) idiom (R/model.py:326-327, R/inference.py:121,
// R/cv_inference.py:105, R/training/trainer_unfreeze.py:1073-1074); the N x M matrix is the
// north_star's superset whose diagonal equals the pairwise scores.
//
//   c_row_norms   one warp per row (both operands in one launch): 1 / max(||x||, 1e-12) and the "some norm is off by > 1e-4" flag
//                 (torch.allclose(norm, 1, atol=1e-4): |norm - 1| <= 1e-4 + 1e-5)
//   c_pairwise    one warp per row: <a_i, b_i> * inv_a[i] * inv_b[i]
//   c_split + c_nxm_tc   the N x M matrix on tcgen05 tensor cores with split-TF32 operands (see below)
#include "stx_common.h"
#include <cuda_bf16.h>
#include <cuda.h>
#include <algorithm>

namespace stx {
namespace {

struct CosWs {            // layout of the workspace
    int    flag_a, flag_b;   // set when the operand has to be re-normalised
    int    pad[62];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float row_sumsq(const float* __restrict__ p, int D, int lane) {
    float acc = 0.0f;
    if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        for (int i = lane; i < D / 4; i += 32) {
            const float4 v = __ldg(p4 + i);
            acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
        }
    } else {
        for (int i = lane; i < D; i += 32) { const float v = __ldg(p + i); acc = fmaf(v, v, acc); }
    }
    return warp_sum(acc);
}

// rows [0, N) are operand a, rows [N, N + M) operand b (one launch for both)
__global__ void __launch_bounds__(256)
c_row_norms(const float* __restrict__ a, const float* __restrict__ b, int N, int M, int D, float* __restrict__ inv_a,
            float* __restrict__ inv_b, CosWs* __restrict__ ws) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N + M) return;
    const bool is_b = row >= N;
    if (is_b) row -= N;
    const float acc = row_sumsq((is_b ? b : a) + (size_t)row * D, D, lane);
    if (lane == 0) {
        const float nrm = sqrtf(acc);
        (is_b ? inv_b : inv_a)[row] = 1.0f / fmaxf(nrm, 1e-12f);
        if (!(fabsf(nrm - 1.0f) <= 1e-4f + 1e-5f)) atomicOr(is_b ? &ws->flag_b : &ws->flag_a, 1);
    }
}

__global__ void __launch_bounds__(256)
c_pairwise(const float* __restrict__ a, const float* __restrict__ b, int N, int D, const float* __restrict__ inv_a,
           const float* __restrict__ inv_b, const CosWs* __restrict__ ws, int always, float* __restrict__ s) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    const float* pa = a + (size_t)row * D;
    const float* pb = b + (size_t)row * D;
    float acc = 0.0f;
    if ((D & 3) == 0 && (((reinterpret_cast<uintptr_t>(pa) | reinterpret_cast<uintptr_t>(pb)) & 15) == 0)) {
        const float4* a4 = reinterpret_cast<const float4*>(pa);
        const float4* b4 = reinterpret_cast<const float4*>(pb);
        for (int i = lane; i < D / 4; i += 32) {
            const float4 u = __ldg(a4 + i), v = __ldg(b4 + i);
            acc = fmaf(u.x, v.x, acc); acc = fmaf(u.y, v.y, acc); acc = fmaf(u.z, v.z, acc); acc = fmaf(u.w, v.w, acc);
        }
    } else {
        for (int i = lane; i < D; i += 32) acc = fmaf(__ldg(pa + i), __ldg(pb + i), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        const float sa = (always || ws->flag_a) ? inv_a[row] : 1.0f;
        const float sb = (always || ws->flag_b) ? inv_b[row] : 1.0f;
        s[row] = acc * sa * sb;
    }
}

// ---- the evaluation-time consumers of the pairwise scores, fused ---------------------------
// One warp per item: s_pos = <aud, pos>, s_neg = <aud, neg> after F.normalize (R/training/trainer_unfreeze.py:561-563,
// 1206-1207), hr = sigmoid(s / temperature) (to_human_readable, :924-939), per-sample 2-way InfoNCE
// CE([s_pos, s_neg] / temperature, 0) = softplus((s_neg - s_pos) / temperature) times an optional alignment factor
// (:722-733); c_pos_neg_loss: mean(per-sample) + gamma * mean(relu(s_neg)) (:735-739), one CTA, fixed summation order.
struct PosNegOut { float *s_pos, *s_neg, *hr_pos, *hr_neg, *per_sample; };

__global__ void __launch_bounds__(256)
c_pos_neg(const float* __restrict__ aud, const float* __restrict__ pos, const float* __restrict__ neg, int B, int D,
          float inv_temperature, const float* __restrict__ align_factor, const PosNegOut o) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* pa = aud + (size_t)row * D;
    const float* pp = pos + (size_t)row * D;
    const float* pn = neg + (size_t)row * D;
    float aa = 0.0f, bb = 0.0f, cc = 0.0f, ab = 0.0f, ac = 0.0f;
    for (int i = lane; i < D; i += 32) {
        const float a = __ldg(pa + i), b = __ldg(pp + i), c = __ldg(pn + i);
        aa = fmaf(a, a, aa); bb = fmaf(b, b, bb); cc = fmaf(c, c, cc);
        ab = fmaf(a, b, ab); ac = fmaf(a, c, ac);
    }
    aa = warp_sum(aa); bb = warp_sum(bb); cc = warp_sum(cc); ab = warp_sum(ab); ac = warp_sum(ac);
    if (lane == 0) {
        const float ia = 1.0f / fmaxf(sqrtf(aa), 1e-12f);
        const float s_pos = ab * ia * (1.0f / fmaxf(sqrtf(bb), 1e-12f));
        const float s_neg = ac * ia * (1.0f / fmaxf(sqrtf(cc), 1e-12f));
        const float z = (s_neg - s_pos) * inv_temperature;
        float per = fmaxf(z, 0.0f) + log1pf(expf(-fabsf(z)));          // softplus, stable
        if (align_factor) per *= align_factor[row];
        o.s_pos[row] = s_pos;
        o.s_neg[row] = s_neg;
        o.hr_pos[row] = 1.0f / (1.0f + expf(-s_pos * inv_temperature));
        o.hr_neg[row] = 1.0f / (1.0f + expf(-s_neg * inv_temperature));
        o.per_sample[row] = per;
    }
}

__global__ void __launch_bounds__(256)
c_pos_neg_loss(const float* __restrict__ per_sample, const float* __restrict__ s_neg, int B, float gamma, float* __restrict__ loss) {
    __shared__ double red[2][256];
    double a = 0.0, r = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) { a += (double)per_sample[i]; r += (double)fmaxf(s_neg[i], 0.0f); }
    red[0][threadIdx.x] = a; red[1][threadIdx.x] = r;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { red[0][threadIdx.x] += red[0][threadIdx.x + s]; red[1][threadIdx.x] += red[1][threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = (float)(red[0][0] / B + (gamma > 0.0f ? (double)gamma * red[1][0] / B : 0.0));
}

// ---- N x M contraction on the 5th-generation tensor cores (tcgen05, accumulators in TMEM) -------------------
//
// float32-grade accuracy from narrow MMAs by operand splitting: x = hi + lo,
//   <a, b> ~= <a_hi, b_hi> + <a_lo, b_hi> + <a_hi, b_lo>
// in one of two operand formats (TcGeom::bf16):
//   TF32 (feature projection): hi = x truncated to TF32 (exactly representable, so the tensor core's own fp32 -> tf32
//        conversion cannot change it), lo = x - hi (exact); the dropped <a_lo, b_lo> is ~2^-22 |a||b|.  K = 8 per MMA.
//   BF16 (cosine scores, round 2): hi = bf16(x), lo = bf16(x - hi), both round-to-nearest: x - hi - lo <= 2^-18 |x|, so
//        the dropped terms are <= 3 * 2^-18 sum|a_i b_i| in the worst case (1.1e-5 for unit rows whose products all have
//        one sign) and ~1e-5 / sqrt(D) in practice: measured <= 2e-6 on unit rows incl. cos = 1 pairs (bar 1e-5).  K = 16
//        per MMA at the same issue rate: half the tensor-pipe time of TF32, half the operand bytes (planes, L2 -> shared
//        memory traffic, NVLink push of the gathered call).  cfg5: 0.125 -> see DESIGN.md.
//
//   c_split    one warp per row: scale by 1/||x|| (or 1), write hi and lo planes with rows zero-padded to a
//              multiple of 32 floats (one 128-byte swizzle row per k-block)
//   c_nxm_tc   persistent: one CTA per SM walks 128 x 128 output tiles (column-major, so CTAs running together
//              share the B tile in L2).  192 threads: warp 0 = TMA producer (cp.async.bulk.tensor, SWIZZLE_128B boxes
//              of 128 rows x 128 B; per k-block the hi AND lo planes of both operands land once, 64 KB per stage,
//              3-stage full/empty mbarrier ring), warp 1 = TMEM allocator + single-thread tcgen05.mma issuer
//              (kind::tf32, M = 128, N = 128, K = 8: per k-block hi.hi, lo.hi and hi.lo from the same stage;
//              tcgen05.commit frees the stage), warps 2-5 = epilogue (tcgen05.ld 32x32b.x32 -> registers -> 128-byte
//              row segments of S).  Two accumulators in TMEM (2 x 128 columns): the epilogue of tile i overlaps the
//              main loop of tile i + 1.
// Stage geometry (A/B: tools/ab_build.py): rows of kRowBytes per k-block in shared memory.  128-byte rows (SWIZZLE_128B) give 3
// stages of 64 KB; 64-byte rows (SWIZZLE_64B) give 7 stages of 32 KB, i.e. 192 KB instead of 128 KB of operands in flight
// behind the stage being consumed -- with bfloat16 operands the kernel is bound by TMA latency x bandwidth, not by the tensor pipe.
#ifndef STX_C_ROW_BYTES
#define STX_C_ROW_BYTES 128
#endif
constexpr int kRowBytes = STX_C_ROW_BYTES;
static_assert(kRowBytes == 128 || kRowBytes == 64, "k-block rows are 128 or 64 bytes");
constexpr int kTM = 128, kTN = 128, kTK = 32, kStages = kRowBytes == 128 ? 3 : 7;
constexpr int kTcThreads = 192;
constexpr int kTileBytes = kTM * kRowBytes;     // one plane tile: 128 rows of one k-block
constexpr unsigned kTmemCols = 512;             // two tiles in flight x two 128 x 128 float accumulators (hi.hi | cross terms)
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6), A = B = TF32 [7,10) [10,13), K-major A and B,
// N >> 3 at [17,23), M >> 4 at [24,29)
constexpr unsigned kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(kTN >> 3) << 17) | ((unsigned)(kTM >> 4) << 24);
// kind::f16 with BF16 operands (format 1 in both fields), F32 accumulator
constexpr unsigned kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(kTN >> 3) << 17) | ((unsigned)(kTM >> 4) << 24);

struct TcSmem {
    unsigned char a_hi[kStages][kTileBytes];    // [128 rows][128 B], 128-byte swizzle, 1024-byte aligned
    unsigned char a_lo[kStages][kTileBytes];
    unsigned char b_hi[kStages][kTileBytes];
    unsigned char b_lo[kStages][kTileBytes];
    unsigned long long full[kStages], empty[kStages], tmem_full[2], tmem_empty[2];
    unsigned tmem_base;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor) of a K-major tile with 128-byte swizzle: start address
// >> 4 at [0,14), leading byte offset (unused for swizzled K-major) = 1 at [16,30), stride byte offset = 1024 B (8 rows)
// >> 4 at [32,46), version 1 at [46,48), layout SWIZZLE_128B = 2 at [61,64)
__device__ __forceinline__ unsigned long long umma_desc(const void* tile, int k_bytes) {
    const unsigned addr = smem_u32(tile) + (unsigned)k_bytes;
    // stride between 8-row groups: 8 x row bytes; layout type 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return (unsigned long long)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((unsigned long long)((8 * kRowBytes) >> 4) << 32) |
           (1ull << 46) | ((unsigned long long)(kRowBytes == 128 ? 2 : 4) << 61);
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdescBf16), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Where c_split writes.  Planes are [hi | lo], each `*_plane_stride` floats apart, rows zero-padded to Dp floats.
// The b planes can go to several destinations at once: on one GPU that is the local workspace; across GPUs every
// destination is the slot of this rank inside a peer's symmetric buffer (NVLink P2P stores), i.e. the all-gather of
// the text embeddings is fused into this kernel, followed by a release of flags[p][flag_index] = epoch on every peer.
constexpr int kMaxWorld = 8;
struct SplitDst {
    float*    a_planes;
    size_t    a_plane_stride;
    float*    b_mc;                 // NVSwitch multicast address of this rank's slot (one multimem.st reaches every
                                    // peer), or nullptr: unicast P2P stores to b_dst[0 .. n_dst)
    float*    b_dst[kMaxWorld];
    size_t    b_plane_stride;
    int       n_dst;
    unsigned* flags[kMaxWorld];     // nullptr entries: no flag protocol (single GPU)
    int       flag_index;
    unsigned  epoch;
    unsigned* counter;              // local: CTAs that have finished their stores
};

// hi / lo planes of both operands (rows [0, N) = a, [N, N + M) = b).  inv == nullptr: normalise unconditionally
// with the norm computed here (no separate norm pass, no flags).  kBf16: the planes are bfloat16 (rows of Dp elements, Dp a
// multiple of 64; plane strides count bfloat16 elements), else float32 holding TF32 values (Dp a multiple of 32).
template <bool kBf16>
__global__ void __launch_bounds__(256)
c_split(const float* __restrict__ a, const float* __restrict__ b, int N, int M, int D, int Dp,
        const float* __restrict__ inv_a, const float* __restrict__ inv_b, const CosWs* __restrict__ ws, const SplitDst dst) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row < N + M) {
        const bool is_b = row >= N;
        if (is_b) row -= N;
        const float* p = (is_b ? b : a) + (size_t)row * D;
        float scale;
        if (inv_a == nullptr) scale = 1.0f / fmaxf(sqrtf(row_sumsq(p, D, lane)), 1e-12f);
        else scale = (is_b ? ws->flag_b : ws->flag_a) ? (is_b ? inv_b : inv_a)[row] : 1.0f;
        const size_t off = (size_t)row * Dp;
        const bool vec_in = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
        if (kBf16) {
            __nv_bfloat16* const a_pl = reinterpret_cast<__nv_bfloat16*>(dst.a_planes);
            for (int i = 8 * lane; i < Dp; i += 256) {                 // Dp is a multiple of 64: whole groups of 8
                float v[8];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (vec_in && i + 4 * h + 4 <= D) {
                        const float4 x = __ldg(reinterpret_cast<const float4*>(p + i + 4 * h));
                        v[4 * h] = x.x * scale; v[4 * h + 1] = x.y * scale; v[4 * h + 2] = x.z * scale; v[4 * h + 3] = x.w * scale;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[4 * h + j] = i + 4 * h + j < D ? __ldg(p + i + 4 * h + j) * scale : 0.0f;
                    }
                }
                unsigned hw[4], lw[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
                    const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * j] - __bfloat162float(h0));
                    const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1));
                    hw[j] = (unsigned)__bfloat16_as_ushort(h0) | ((unsigned)__bfloat16_as_ushort(h1) << 16);
                    lw[j] = (unsigned)__bfloat16_as_ushort(l0) | ((unsigned)__bfloat16_as_ushort(l1) << 16);
                }
                const uint4 H = make_uint4(hw[0], hw[1], hw[2], hw[3]), L = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                if (!is_b) {
                    *reinterpret_cast<uint4*>(a_pl + off + i) = H;
                    *reinterpret_cast<uint4*>(a_pl + dst.a_plane_stride + off + i) = L;
                } else if (dst.b_mc) {
                    __nv_bfloat16* const mc = reinterpret_cast<__nv_bfloat16*>(dst.b_mc);
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                                 ::"l"(mc + off + i), "f"(__uint_as_float(H.x)), "f"(__uint_as_float(H.y)),
                                   "f"(__uint_as_float(H.z)), "f"(__uint_as_float(H.w)) : "memory");
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                                 ::"l"(mc + dst.b_plane_stride + off + i), "f"(__uint_as_float(L.x)), "f"(__uint_as_float(L.y)),
                                   "f"(__uint_as_float(L.z)), "f"(__uint_as_float(L.w)) : "memory");
                } else {
#pragma unroll 1
                    for (int d = 0; d < dst.n_dst; ++d) {
                        __nv_bfloat16* const bp = reinterpret_cast<__nv_bfloat16*>(dst.b_dst[d]);
                        *reinterpret_cast<uint4*>(bp + off + i) = H;
                        *reinterpret_cast<uint4*>(bp + dst.b_plane_stride + off + i) = L;
                    }
                }
            }
        } else
        for (int i = 4 * lane; i < Dp; i += 128) {                     // Dp is a multiple of 32: whole float4s
            float v[4];
            if (vec_in && i + 4 <= D) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(p + i));
                v[0] = x.x * scale; v[1] = x.y * scale; v[2] = x.z * scale; v[3] = x.w * scale;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = i + j < D ? __ldg(p + i + j) * scale : 0.0f;
            }
            float4 h, l;
            h.x = __int_as_float(__float_as_int(v[0]) & 0xffffe000);     // TF32: 10 explicit mantissa bits
            h.y = __int_as_float(__float_as_int(v[1]) & 0xffffe000);
            h.z = __int_as_float(__float_as_int(v[2]) & 0xffffe000);
            h.w = __int_as_float(__float_as_int(v[3]) & 0xffffe000);
            l = make_float4(v[0] - h.x, v[1] - h.y, v[2] - h.z, v[3] - h.w);
            if (!is_b) {
                *reinterpret_cast<float4*>(dst.a_planes + off + i) = h;
                *reinterpret_cast<float4*>(dst.a_planes + dst.a_plane_stride + off + i) = l;
            } else if (dst.b_mc) {
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                             ::"l"(dst.b_mc + off + i), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                             ::"l"(dst.b_mc + dst.b_plane_stride + off + i), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
            } else {
#pragma unroll 1
                for (int d = 0; d < dst.n_dst; ++d) {
                    *reinterpret_cast<float4*>(dst.b_dst[d] + off + i) = h;
                    *reinterpret_cast<float4*>(dst.b_dst[d] + dst.b_plane_stride + off + i) = l;
                }
            }
        }
    }
    if (dst.counter) {
        // publish: every CTA's stores are fenced system-wide, the last CTA releases the flag on every peer
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned done = atomicAdd(dst.counter, 1u);
            if (done == gridDim.x - 1) {
                __threadfence_system();
                for (int d = 0; d < dst.n_dst; ++d)
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst.flags[d] + dst.flag_index), "r"(dst.epoch) : "memory");
                *dst.counter = 0;
            }
        }
    }
}

// The two-set protocol of stx_cosine_nxm_gathered has no barrier between calls: it is safe only if, in EVERY call, every rank
// acquires the flag of EVERY peer before its stream moves on (a peer's push of call e + 2 reuses the planes of call e).  The
// GEMM's TMA producer acquires the flags of the slots it reads; a rank that has nothing to compute (no local rows, or no
// columns at all), and the slots that hold no rows, are covered by this one-warp kernel instead.
__global__ void c_acquire_flags(const unsigned* __restrict__ flags, int world, unsigned epoch) {
    const int r = threadIdx.x;
    if (r < world) {
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
            if ((int)(v - epoch) < 0) __nanosleep(64);
        } while ((int)(v - epoch) < 0);
    }
}

// ---- encoder input stage: LayerNorm(in_dim) + Linear(in_dim -> out_dim) -------------------------
// Wav2Vec2BertFeatureProjection.forward (TF/models/wav2vec2_bert/modeling_wav2vec2_bert.py:118-130), the first thing
// the speech encoder does with input_features (:1016).  p_ln_split: one warp per row, torch.nn.LayerNorm semantics
// (biased variance, eps inside the sqrt), optional copy of the normalised row (the module's second return value), and
// the TF32 hi / lo planes of it; p_split: planes of the weight matrix; the contraction is c_nxm_tc with a bias epilogue.
__global__ void __launch_bounds__(256)
p_ln_split(const float* __restrict__ x, int rows, int D, int Dp, const float* __restrict__ gamma, const float* __restrict__ beta,
           float eps, float* __restrict__ norm_out, float* __restrict__ planes, size_t plane_stride) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* p = x + (size_t)row * D;
    float s = 0.0f;
    for (int i = lane; i < D; i += 32) s += __ldg(p + i);
    const float mean = warp_sum(s) / (float)D;
    float q = 0.0f;
    for (int i = lane; i < D; i += 32) { const float d = __ldg(p + i) - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    const size_t off = (size_t)row * Dp;
    for (int i = lane; i < Dp; i += 32) {
        float v = 0.0f;
        if (i < D) {
            v = (__ldg(p + i) - mean) * rstd * __ldg(gamma + i) + __ldg(beta + i);
            if (norm_out) norm_out[(size_t)row * D + i] = v;
        }
        const float h = __int_as_float(__float_as_int(v) & 0xffffe000);
        planes[off + i] = h;
        planes[plane_stride + off + i] = v - h;
    }
}

__global__ void __launch_bounds__(256)
p_split(const float* __restrict__ w, int rows, int D, int Dp, float* __restrict__ planes, size_t plane_stride) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const size_t off = (size_t)row * Dp;
    for (int i = lane; i < Dp; i += 32) {
        const float v = i < D ? __ldg(w + (size_t)row * D + i) : 0.0f;
        const float h = __int_as_float(__float_as_int(v) & 0xffffe000);
        planes[off + i] = h;
        planes[plane_stride + off + i] = v - h;
    }
}

// Geometry of one launch of c_nxm_tc.  B lives in `world` slots of [hi | lo] planes of b_plane_rows rows each (one slot
// per source rank; one slot on a single GPU); column tiles are visited in `slot[]` order (own rank first).
struct TcGeom {
    int n_rows;                     // valid rows of A
    int a_plane_rows;               // rows between A's hi and lo planes
    int b_plane_rows;               // rows per plane of a B slot
    int world;
    int tiles_start[kMaxWorld + 1]; // prefix sums of the column tiles per slot, in visiting order
    int slot[kMaxWorld];            // visiting order -> slot
    int m_count[kMaxWorld];         // valid rows (output columns) of a slot
    int col_start[kMaxWorld];       // first output column of a slot
    int ldS;                        // row pitch of S (total columns)
    const unsigned* flags;          // flags[slot] reaches `epoch` when the slot's planes have landed (nullptr: local data)
    unsigned epoch;
    const float* bias;              // nullptr, or one value per output column added in the epilogue (feature projection)
    int bf16;                       // operand planes are bfloat16 (64 elements per 128-byte k-block row), else TF32 in float32 (32)
    // retrieval mode (top-k): the epilogue does not write S; every (row, column tile) leaves its kTopK best scores and their
    // column indices (descending; ties: lower column first) in cand_val / cand_idx [row][column tile][kTopK]
    float* cand_val;
    int*   cand_idx;
};
constexpr int kTopK = 8;            // candidates kept per (row, column tile); stx_cosine_topk serves k <= kTopK

__global__ void __launch_bounds__(kTcThreads, 1)
c_nxm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
         const __grid_constant__ TcGeom g, int kb_per_pass, int n_col_tiles, float* __restrict__ S) {
    extern __shared__ unsigned char smem_raw[];
    TcSmem& sm = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int N = g.n_rows, M = g.ldS;
    const int n_row_tiles = (N + kTM - 1) / kTM;
    const int num_tiles = n_row_tiles * n_col_tiles;

    // tile t -> (row tile, column tile, slot geometry); column-major: tiles that run together share their B tile
    struct Tile { int m0, slot, j0, n0, n_valid, b_row_hi; };
    auto tile_of = [&](int t) {
        Tile ti;
        const int ct = t / n_row_tiles;
        ti.m0 = (t - ct * n_row_tiles) * kTM;
        int ord = 0;
        while (ord + 1 < g.world && ct >= g.tiles_start[ord + 1]) ++ord;
        ti.slot = g.slot[ord];
        ti.j0 = (ct - g.tiles_start[ord]) * kTN;                        // first row of the tile inside its slot
        ti.n0 = g.col_start[ti.slot] + ti.j0;                           // first output column
        ti.n_valid = min(kTN, g.m_count[ti.slot] - ti.j0);
        ti.b_row_hi = (ti.slot * 2) * g.b_plane_rows + ti.j0;
        return ti;
    };

    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(&sm.tmem_full[a], 1); mbar_init(&sm.tmem_empty[a], 4); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sm.tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;                              // k-blocks issued so far, over all of this CTA's tiles
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const Tile ti = tile_of(t);
                if (g.flags) {
                    // the slot's planes are written by its source rank over NVLink: acquire its flag, then order the
                    // generic-proxy view before the async-proxy (TMA) reads
                    unsigned v;
                    do {
                        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(g.flags + ti.slot) : "memory");
                        if ((int)(v - g.epoch) < 0) __nanosleep(64);
                    } while ((int)(v - g.epoch) < 0);
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                const int kelems = g.bf16 ? kRowBytes / 2 : kRowBytes / 4;      // elements per k-block row
                for (int kb = 0; kb < kb_per_pass; ++kb, ++it) {
                    const int s = it % kStages, round = it / kStages;
                    if (round > 0) mbar_wait(&sm.empty[s], (unsigned)(round - 1) & 1u);
                    mbar_expect_tx(&sm.full[s], 4 * kTileBytes);
                    tma_load_2d(sm.a_hi[s], &map_a, kb * kelems, ti.m0, &sm.full[s]);
                    tma_load_2d(sm.a_lo[s], &map_a, kb * kelems, ti.m0 + g.a_plane_rows, &sm.full[s]);
                    tma_load_2d(sm.b_hi[s], &map_b, kb * kelems, ti.b_row_hi, &sm.full[s]);
                    tma_load_2d(sm.b_lo[s], &map_b, kb * kelems, ti.b_row_hi + g.b_plane_rows, &sm.full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            int it = 0, n = 0;
            // Descriptors of stage 0, built ONCE: the planes of a stage are kTileBytes apart and a K step is 32 bytes, so the
            // descriptor of (stage s, step k) is the base plus s * (kTileBytes >> 4) + 2 k in its 16-byte address field.  (Built
            // per MMA -- a generic-to-shared conversion and half a dozen dependent integer operations each, on a single thread --
            // the issue loop took ~170 cycles per MMA and the tensor pipe waited for it: 0.097 -> see DESIGN.md.)
            const unsigned long long ah0 = umma_desc(sm.a_hi[0], 0), al0 = umma_desc(sm.a_lo[0], 0);
            const unsigned long long bh0 = umma_desc(sm.b_hi[0], 0), bl0 = umma_desc(sm.b_lo[0], 0);
            const bool bf = g.bf16 != 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++n) {
                const int acc = n & 1;
                if (n >= 2) mbar_wait(&sm.tmem_empty[acc], (unsigned)(n / 2 - 1) & 1u);     // the epilogue has drained it
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // TWO accumulators per tile: the tensor core truncates (rounds toward zero) after every MMA, which biases a sum of
                // like-signed products by about -1.8e-8 * D * score when all 3 D / 8 instructions of a tile land on one running
                // sum (-1.9e-5 at D = 1024, score 1: above the 1e-5 bar exactly for the well-matched pairs the scores are for).
                // A_hi B_hi carries the magnitude, the two cross products are 2^-11 of it: kept apart, only D / 8 truncations
                // touch the large sum (bias / 3) and the small one loses nothing that matters; the epilogue adds them in float32.
                const unsigned d = tmem + (unsigned)(acc * 2 * kTN), dx = d + (unsigned)kTN;
                for (int kb = 0; kb < kb_per_pass; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(&sm.full[s], (unsigned)(it / kStages) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned long long so = (unsigned long long)(s * (kTileBytes >> 4));
                    const unsigned long long ah = ah0 + so, al = al0 + so, bh = bh0 + so, bl = bl0 + so;
                    if (bf) {                       // 32 bytes per K step in both formats: 16 bfloat16 or 8 TF32
#pragma unroll
                        for (int k = 0; k < kRowBytes / 32; ++k) {
                            umma_bf16(d, ah + 2 * k, bh + 2 * k, (kb | k) != 0);
                            umma_bf16(dx, al + 2 * k, bh + 2 * k, (kb | k) != 0);
                            umma_bf16(dx, ah + 2 * k, bl + 2 * k, 1);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < kRowBytes / 32; ++k) {
                            umma_tf32(d, ah + 2 * k, bh + 2 * k, (kb | k) != 0);
                            umma_tf32(dx, al + 2 * k, bh + 2 * k, (kb | k) != 0);
                            umma_tf32(dx, ah + 2 * k, bl + 2 * k, 1);
                        }
                    }
                    umma_commit(&sm.empty[s]);      // the stage is free once these MMAs have read it
                }
                umma_commit(&sm.tmem_full[acc]);    // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp & 3;                     // a warp can only read its own quarter of the TMEM lanes
        const bool vec = (M & 3) == 0 && (reinterpret_cast<uintptr_t>(S) & 15) == 0;
        int n = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++n) {
            const Tile ti = tile_of(t);
            const int acc = n & 1;
            mbar_wait(&sm.tmem_full[acc], (unsigned)(n / 2) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = ti.m0 + q * 32 + lane;
            float bv[kTopK];
            int bi[kTopK];
#pragma unroll
            for (int i = 0; i < kTopK; ++i) { bv[i] = __int_as_float(0xff800000); bi[i] = 0x7fffffff; }
#pragma unroll 1
            for (int c0 = 0; c0 < kTN; c0 += 32) {
                unsigned r[32], rx[32];
                const unsigned taddr = tmem + ((unsigned)(q * 32) << 16) + (unsigned)(acc * 2 * kTN + c0);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                               "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(taddr));
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(rx[0]), "=r"(rx[1]), "=r"(rx[2]), "=r"(rx[3]), "=r"(rx[4]), "=r"(rx[5]), "=r"(rx[6]), "=r"(rx[7]),
                               "=r"(rx[8]), "=r"(rx[9]), "=r"(rx[10]), "=r"(rx[11]), "=r"(rx[12]), "=r"(rx[13]), "=r"(rx[14]), "=r"(rx[15]),
                               "=r"(rx[16]), "=r"(rx[17]), "=r"(rx[18]), "=r"(rx[19]), "=r"(rx[20]), "=r"(rx[21]), "=r"(rx[22]), "=r"(rx[23]),
                               "=r"(rx[24]), "=r"(rx[25]), "=r"(rx[26]), "=r"(rx[27]), "=r"(rx[28]), "=r"(rx[29]), "=r"(rx[30]), "=r"(rx[31])
                             : "r"(taddr + (unsigned)kTN));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(rx[j]));
                if (c0 >= ti.n_valid) continue;
                if (g.bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < ti.n_valid) r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldg(g.bias + ti.n0 + c0 + j));
                }
                if (g.cand_val) {
                    // a score enters the sorted list only if it beats the current k-th best: after the first few columns
                    // that is rare, so the insertion chain is off the common path
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = __uint_as_float(r[j]);
                        if (c0 + j < ti.n_valid && v > bv[kTopK - 1]) {
                            float cv = v;
                            int ci = ti.n0 + c0 + j;
#pragma unroll
                            for (int i = 0; i < kTopK; ++i) {
                                if (cv > bv[i]) {            // strict: an equal score keeps the earlier (lower) column ahead
                                    const float tv = bv[i]; const int tix = bi[i];
                                    bv[i] = cv; bi[i] = ci; cv = tv; ci = tix;
                                }
                            }
                        }
                    }
                } else if (row < N) {
                    float* dst = S + (size_t)row * M + ti.n0 + c0;
                    if (vec && (ti.n0 & 3) == 0 && c0 + 32 <= ti.n_valid) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            reinterpret_cast<float4*>(dst)[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c0 + j < ti.n_valid) dst[j] = __uint_as_float(r[j]);
                    }
                }
            }
            if (g.cand_val && row < N) {
                const size_t base = ((size_t)row * n_col_tiles + (size_t)(ti.n0 / kTN)) * kTopK;
#pragma unroll
                for (int i = 0; i < kTopK; ++i) { g.cand_val[base + i] = bv[i]; g.cand_idx[base + i] = bi[i]; }
            }
            // this warp is done with the accumulator: one arrival per epilogue warp frees it for tile n + 2
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sm.tmem_empty[acc])) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// Top-k of a row from its per-column-tile candidate lists (c_nxm_tc in retrieval mode).  One warp per row: lane L first
// merges the lists of tiles L, L + 32, ... into one sorted list of kTopK, then k rounds of a warp-wide arg-max over the
// lanes' heads (score descending, column ascending on ties) pop the winners in order.
__device__ __forceinline__ bool topk_before(float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); }
__global__ void __launch_bounds__(256)
c_topk_merge(const float* __restrict__ cand_val, const int* __restrict__ cand_idx, int N, int n_col_tiles, int k,
             float* __restrict__ out_val, int* __restrict__ out_idx) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    float lv[kTopK];
    int li[kTopK];
#pragma unroll
    for (int i = 0; i < kTopK; ++i) { lv[i] = __int_as_float(0xff800000); li[i] = 0x7fffffff; }
    for (int t = lane; t < n_col_tiles; t += 32) {
        const size_t base = ((size_t)row * n_col_tiles + t) * kTopK;
#pragma unroll
        for (int j = 0; j < kTopK; ++j) {
            float cv = __ldg(cand_val + base + j);
            int ci = __ldg(cand_idx + base + j);
#pragma unroll
            for (int i = 0; i < kTopK; ++i) {
                if (topk_before(cv, ci, lv[i], li[i])) {
                    const float tv = lv[i]; const int ti = li[i];
                    lv[i] = cv; li[i] = ci; cv = tv; ci = ti;
                }
            }
        }
    }
    for (int r = 0; r < k; ++r) {
        float bv = lv[0];
        int bi = li[0], bl = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o), ol = __shfl_xor_sync(0xffffffffu, bl, o);
            if (topk_before(ov, oi, bv, bi) || (ov == bv && oi == bi && ol < bl)) { bv = ov; bi = oi; bl = ol; }
        }
        if (lane == 0) { out_val[(size_t)row * k + r] = bv; out_idx[(size_t)row * k + r] = bi == 0x7fffffff ? -1 : bi; }
        if (lane == bl) {                            // the winner pops its head
#pragma unroll
            for (int i = 0; i + 1 < kTopK; ++i) { lv[i] = lv[i + 1]; li[i] = li[i + 1]; }
            lv[kTopK - 1] = __int_as_float(0xff800000); li[kTopK - 1] = 0x7fffffff;
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_map(CUtensorMap* map, const void* base, int rows, int Dp, int box_rows, bool bf16) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        STX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available"); return STX_EDEVICE; }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)Dp, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)Dp * (bf16 ? 2 : 4)};
    const cuuint32_t box[2] = {(cuuint32_t)(bf16 ? kRowBytes / 2 : kRowBytes / 4), (cuuint32_t)box_rows};   // kRowBytes per box row
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, kRowBytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return STX_EINVAL; }
    return 0;
}

inline int padded_d(int D) { return (D + kTK - 1) / kTK * kTK; }
inline int padded_d64(int D) { return (D + 2 * kTK - 1) / (2 * kTK) * (2 * kTK); }
// operand format of the cosine contractions: bfloat16 hi / lo planes unless STX_COSINE_TF32=1 (A/B runs; every rank of a
// gathered call must agree)
inline bool cosine_bf16() {
    static const bool v = [] { const char* e = std::getenv("STX_COSINE_TF32"); return !(e && e[0] == '1'); }();
    return v;
}
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

int launch_gemm(const void* a_planes, int a_rows_total, const void* b_planes, int b_rows_total, int Dp, const TcGeom& g,
                int n_col_tiles, float* d_S, cudaStream_t st) {
    CUtensorMap ma, mb;
    if (int rc = make_map(&ma, a_planes, a_rows_total, Dp, kTM, g.bf16 != 0)) return rc;
    if (int rc = make_map(&mb, b_planes, b_rows_total, Dp, kTN, g.bf16 != 0)) return rc;
    static bool attr_set[64] = {false};
    int dev = 0;
    STX_CUDA(cudaGetDevice(&dev));
    const int smem_bytes = (int)sizeof(TcSmem) + 1024;
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        STX_CUDA(cudaFuncSetAttribute(c_nxm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_set[dev] = true;
    }
    static int sms[64] = {0};
    if (dev >= 0 && dev < 64 && !sms[dev]) STX_CUDA(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
    const long long tiles = (long long)n_col_tiles * ((g.n_rows + kTM - 1) / kTM);
    const int ctas = (int)std::min<long long>(tiles, (dev >= 0 && dev < 64 && sms[dev] > 0) ? sms[dev] : 148);
    STX_LAUNCH(c_nxm_tc, dim3(ctas), dim3(kTcThreads), smem_bytes, st, ma, mb, g, Dp / (g.bf16 ? kRowBytes / 2 : kRowBytes / 4),
               n_col_tiles, d_S);
    return 0;
}

// layout of a rank's symmetric buffer for the gathered N x M call: [world][hi | lo][m_cap][Dp] floats, then flags
inline size_t symm_planes_bytes(int world, int m_cap, int Dp) { return align256(size_t(world) * 2 * m_cap * Dp * sizeof(float)); }

}  // namespace

// Linear(in_dim -> out_dim) of rows whose hi / lo TF32 planes are already in a_planes ([2][rows][Dp], Dp = in_dim rounded up
// to 32): splits the weight into b_planes and runs the tensor-core contraction with the bias epilogue (used by the fused
// front end, stx_fbank_k_projection in fbank_k.cu; stx_feature_projection above is the same sequence after p_ln_split).
int project_from_planes(const float* a_planes, int rows, int in_dim, const float* d_weight, const float* d_bias, int out_dim,
                        float* b_planes, float* d_hidden, cudaStream_t st) {
    const int Dp = padded_d(in_dim);
    STX_LAUNCH(p_split, dim3((out_dim + 7) / 8), dim3(256), 0, st, d_weight, out_dim, in_dim, Dp, b_planes, size_t(out_dim) * Dp);
    TcGeom g = {};
    g.n_rows = rows;  g.a_plane_rows = rows;  g.b_plane_rows = out_dim;  g.world = 1;
    g.tiles_start[0] = 0;  g.tiles_start[1] = (out_dim + kTN - 1) / kTN;
    g.slot[0] = 0;  g.m_count[0] = out_dim;  g.col_start[0] = 0;  g.ldS = out_dim;
    g.bias = d_bias;
    return launch_gemm(a_planes, 2 * rows, b_planes, 2 * out_dim, Dp, g, g.tiles_start[1], d_hidden, st);
}
}  // namespace stx

extern "C" {

int stx_cosine_workspace(int N, int M, int D, size_t* bytes) {
    using namespace stx;
    if (N < 0 || M < 0 || D < 0 || !bytes) { set_error("stx_cosine_workspace: bad argument"); return STX_EINVAL; }
    const size_t Dp = size_t(padded_d(D));
    *bytes = align256(sizeof(CosWs)) + align256(size_t(N) * sizeof(float)) + align256(size_t(M) * sizeof(float)) +
             2 * align256(size_t(N) * Dp * sizeof(float)) + 2 * align256(size_t(M) * Dp * sizeof(float));
    return 0;
}

static int cosine_prepare(const float* d_a, const float* d_b, int N, int M, int D, void* d_ws, size_t ws_bytes,
                          cudaStream_t st, stx::CosWs** ws, float** inv_a, float** inv_b, bool with_norms = true) {
    using namespace stx;
    size_t need = 0;
    stx_cosine_workspace(N, M, D, &need);
    if (ws_bytes < need) { set_error("cosine: workspace %zu < %zu bytes", ws_bytes, need); return STX_ENOSPACE; }
    char* base = static_cast<char*>(d_ws);
    *ws = reinterpret_cast<CosWs*>(base);
    *inv_a = reinterpret_cast<float*>(base + align256(sizeof(CosWs)));
    *inv_b = reinterpret_cast<float*>(base + align256(sizeof(CosWs)) + align256(size_t(N) * sizeof(float)));
    if (with_norms) {
        STX_CUDA(cudaMemsetAsync(*ws, 0, sizeof(CosWs), st));
        STX_LAUNCH(c_row_norms, dim3((N + M + 7) / 8), dim3(256), 0, st, d_a, d_b, N, M, D, *inv_a, *inv_b, *ws);
    }
    return 0;
}

int stx_cosine_pairwise(const float* d_a, const float* d_b, int N, int D, int always_normalize, float* d_s,
                        void* d_ws, size_t ws_bytes, void* stream) {
    using namespace stx;
    if (N < 0 || D <= 0) { set_error("stx_cosine_pairwise: need N >= 0, D > 0"); return STX_EINVAL; }
    if (N == 0) return 0;
    if (!d_a || !d_b || !d_s || !d_ws) { set_error("stx_cosine_pairwise: null pointer"); return STX_EINVAL; }
    if (int rc = check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CosWs* ws; float *inv_a, *inv_b;
    if (int rc = cosine_prepare(d_a, d_b, N, N, D, d_ws, ws_bytes, st, &ws, &inv_a, &inv_b)) return rc;
    STX_LAUNCH(c_pairwise, dim3((N + 7) / 8), dim3(256), 0, st, d_a, d_b, N, D, inv_a, inv_b, ws, always_normalize, d_s);
    return 0;
}

int stx_cosine_nxm(const float* d_a, const float* d_b, int N, int M, int D, int always_normalize, float* d_S,
                   void* d_ws, size_t ws_bytes, void* stream) {
    using namespace stx;
    if (N < 0 || M < 0 || D <= 0) { set_error("stx_cosine_nxm: need N, M >= 0, D > 0"); return STX_EINVAL; }
    if (N == 0 || M == 0) return 0;
    if (!d_a || !d_b || !d_S || !d_ws) { set_error("stx_cosine_nxm: null pointer"); return STX_EINVAL; }
    if (int rc = check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CosWs* ws; float *inv_a, *inv_b;
    if (int rc = cosine_prepare(d_a, d_b, N, M, D, d_ws, ws_bytes, st, &ws, &inv_a, &inv_b, !always_normalize)) return rc;
    // (the workspace is sized for float32 planes of padded_d(D) columns; bfloat16 planes of padded_d64(D) columns never need more)
    const bool bf = cosine_bf16();
    const int Dp = bf ? padded_d64(D) : padded_d(D);
    char* p = reinterpret_cast<char*>(inv_b) + align256(size_t(M) * sizeof(float));
    float* a_planes = reinterpret_cast<float*>(p);  p += 2 * align256(size_t(N) * padded_d(D) * sizeof(float));
    float* b_planes = reinterpret_cast<float*>(p);
    SplitDst dst = {};
    dst.a_planes = a_planes;  dst.a_plane_stride = size_t(N) * Dp;
    dst.b_dst[0] = b_planes;  dst.b_plane_stride = size_t(M) * Dp;  dst.n_dst = 1;
    if (bf) STX_LAUNCH(c_split<true>, dim3((N + M + 7) / 8), dim3(256), 0, st, d_a, d_b, N, M, D, Dp,
                       always_normalize ? nullptr : inv_a, always_normalize ? nullptr : inv_b, ws, dst);
    else STX_LAUNCH(c_split<false>, dim3((N + M + 7) / 8), dim3(256), 0, st, d_a, d_b, N, M, D, Dp,
                    always_normalize ? nullptr : inv_a, always_normalize ? nullptr : inv_b, ws, dst);
    TcGeom g = {};
    g.n_rows = N;  g.a_plane_rows = N;  g.b_plane_rows = M;  g.world = 1;  g.bf16 = bf;
    g.tiles_start[0] = 0;  g.tiles_start[1] = (M + kTN - 1) / kTN;
    g.slot[0] = 0;  g.m_count[0] = M;  g.col_start[0] = 0;  g.ldS = M;
    return launch_gemm(a_planes, 2 * N, b_planes, 2 * M, Dp, g, g.tiles_start[1], d_S, st);
}

int stx_cosine_topk_workspace(int N, int M, int D, size_t* bytes) {
    using namespace stx;
    size_t base = 0;
    if (int rc = stx_cosine_workspace(N, M, D, &base)) return rc;
    const size_t tiles = size_t((M + kTN - 1) / kTN);
    *bytes = base + align256(size_t(N) * tiles * kTopK * sizeof(float)) + align256(size_t(N) * tiles * kTopK * sizeof(int));
    return 0;
}

int stx_cosine_topk(const float* d_a, const float* d_b, int N, int M, int D, int always_normalize, int k, float* d_val,
                    int32_t* d_idx, void* d_ws, size_t ws_bytes, void* stream) {
    using namespace stx;
    if (N < 0 || M < 0 || D <= 0) { set_error("stx_cosine_topk: need N, M >= 0, D > 0"); return STX_EINVAL; }
    if (k < 1 || k > kTopK) { set_error("stx_cosine_topk: k = %d outside 1..%d", k, kTopK); return STX_EINVAL; }
    if (N == 0) return 0;
    if (M == 0) { set_error("stx_cosine_topk: M = 0 (nothing to retrieve from)"); return STX_EINVAL; }
    if (!d_a || !d_b || !d_val || !d_idx || !d_ws) { set_error("stx_cosine_topk: null pointer"); return STX_EINVAL; }
    if (int rc = check_device()) return rc;
    size_t need = 0, base = 0;
    stx_cosine_topk_workspace(N, M, D, &need);
    stx_cosine_workspace(N, M, D, &base);
    if (ws_bytes < need) { set_error("stx_cosine_topk: workspace %zu < %zu bytes", ws_bytes, need); return STX_ENOSPACE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CosWs* ws; float *inv_a, *inv_b;
    if (int rc = cosine_prepare(d_a, d_b, N, M, D, d_ws, base, st, &ws, &inv_a, &inv_b, !always_normalize)) return rc;
    const bool bf = cosine_bf16();
    const int Dp = bf ? padded_d64(D) : padded_d(D);
    char* p = reinterpret_cast<char*>(inv_b) + align256(size_t(M) * sizeof(float));
    float* a_planes = reinterpret_cast<float*>(p);  p += 2 * align256(size_t(N) * padded_d(D) * sizeof(float));
    float* b_planes = reinterpret_cast<float*>(p);
    const int tiles = (M + kTN - 1) / kTN;
    char* c = static_cast<char*>(d_ws) + base;
    float* cand_val = reinterpret_cast<float*>(c);  c += align256(size_t(N) * tiles * kTopK * sizeof(float));
    int* cand_idx = reinterpret_cast<int*>(c);
    SplitDst dst = {};
    dst.a_planes = a_planes;  dst.a_plane_stride = size_t(N) * Dp;
    dst.b_dst[0] = b_planes;  dst.b_plane_stride = size_t(M) * Dp;  dst.n_dst = 1;
    if (bf) STX_LAUNCH(c_split<true>, dim3((N + M + 7) / 8), dim3(256), 0, st, d_a, d_b, N, M, D, Dp,
                       always_normalize ? nullptr : inv_a, always_normalize ? nullptr : inv_b, ws, dst);
    else STX_LAUNCH(c_split<false>, dim3((N + M + 7) / 8), dim3(256), 0, st, d_a, d_b, N, M, D, Dp,
                    always_normalize ? nullptr : inv_a, always_normalize ? nullptr : inv_b, ws, dst);
    TcGeom g = {};
    g.n_rows = N;  g.a_plane_rows = N;  g.b_plane_rows = M;  g.world = 1;  g.bf16 = bf;
    g.tiles_start[0] = 0;  g.tiles_start[1] = tiles;
    g.slot[0] = 0;  g.m_count[0] = M;  g.col_start[0] = 0;  g.ldS = M;
    g.cand_val = cand_val;  g.cand_idx = cand_idx;
    if (int rc = launch_gemm(a_planes, 2 * N, b_planes, 2 * M, Dp, g, tiles, nullptr, st)) return rc;
    STX_LAUNCH(c_topk_merge, dim3((N + 7) / 8), dim3(256), 0, st, cand_val, cand_idx, N, tiles, k, d_val, d_idx);
    return 0;
}

int stx_score_pos_neg(const float* d_aud, const float* d_pos, const float* d_neg, int B, int D, float temperature,
                      float corrupt_gamma, const float* d_align_factor, float* d_s_pos, float* d_s_neg, float* d_hr_pos,
                      float* d_hr_neg, float* d_per_sample, float* d_loss, void* stream) {
    using namespace stx;
    if (B < 0 || D <= 0 || !(temperature > 0.0f)) { set_error("stx_score_pos_neg: need B >= 0, D > 0, temperature > 0"); return STX_EINVAL; }
    if (B == 0) return 0;
    if (!d_aud || !d_pos || !d_neg || !d_s_pos || !d_s_neg || !d_hr_pos || !d_hr_neg || !d_per_sample || !d_loss) {
        set_error("stx_score_pos_neg: null pointer");
        return STX_EINVAL;
    }
    if (int rc = check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const PosNegOut o = {d_s_pos, d_s_neg, d_hr_pos, d_hr_neg, d_per_sample};
    STX_LAUNCH(c_pos_neg, dim3((B + 7) / 8), dim3(256), 0, st, d_aud, d_pos, d_neg, B, D, 1.0f / temperature, d_align_factor, o);
    STX_LAUNCH(c_pos_neg_loss, dim3(1), dim3(256), 0, st, d_per_sample, d_s_neg, B, corrupt_gamma, d_loss);
    return 0;
}

int stx_feature_projection_workspace(int rows, int in_dim, int out_dim, size_t* bytes) {
    using namespace stx;
    if (rows < 0 || in_dim <= 0 || out_dim <= 0 || !bytes) { set_error("stx_feature_projection_workspace: bad argument"); return STX_EINVAL; }
    const size_t Dp = size_t(padded_d(in_dim));
    *bytes = 2 * align256(size_t(rows) * Dp * sizeof(float)) + 2 * align256(size_t(out_dim) * Dp * sizeof(float));
    return 0;
}

int stx_feature_projection(const float* d_x, const float* d_ln_weight, const float* d_ln_bias, float eps, const float* d_weight,
                           const float* d_bias, int rows, int in_dim, int out_dim, float* d_hidden, float* d_norm,
                           void* d_ws, size_t ws_bytes, void* stream) {
    using namespace stx;
    if (rows < 0 || in_dim <= 0 || out_dim <= 0) { set_error("stx_feature_projection: need rows >= 0, in_dim, out_dim > 0"); return STX_EINVAL; }
    if (rows == 0) return 0;
    if (!d_x || !d_ln_weight || !d_ln_bias || !d_weight || !d_hidden || !d_ws) { set_error("stx_feature_projection: null pointer"); return STX_EINVAL; }
    size_t need = 0;
    stx_feature_projection_workspace(rows, in_dim, out_dim, &need);
    if (ws_bytes < need) { set_error("stx_feature_projection: workspace %zu < %zu bytes", ws_bytes, need); return STX_ENOSPACE; }
    if (int rc = check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Dp = padded_d(in_dim);
    float* a_planes = static_cast<float*>(d_ws);
    float* b_planes = reinterpret_cast<float*>(static_cast<char*>(d_ws) + 2 * align256(size_t(rows) * Dp * sizeof(float)));
    STX_LAUNCH(p_ln_split, dim3((rows + 7) / 8), dim3(256), 0, st, d_x, rows, in_dim, Dp, d_ln_weight, d_ln_bias, eps, d_norm,
               a_planes, size_t(rows) * Dp);
    STX_LAUNCH(p_split, dim3((out_dim + 7) / 8), dim3(256), 0, st, d_weight, out_dim, in_dim, Dp, b_planes, size_t(out_dim) * Dp);
    TcGeom g = {};
    g.n_rows = rows;  g.a_plane_rows = rows;  g.b_plane_rows = out_dim;  g.world = 1;
    g.tiles_start[0] = 0;  g.tiles_start[1] = (out_dim + kTN - 1) / kTN;
    g.slot[0] = 0;  g.m_count[0] = out_dim;  g.col_start[0] = 0;  g.ldS = out_dim;
    g.bias = d_bias;
    return launch_gemm(a_planes, 2 * rows, b_planes, 2 * out_dim, Dp, g, g.tiles_start[1], d_hidden, st);
}

int stx_cosine_gather_sizes(int n_local, int m_cap, int world, int D, size_t* ws_bytes, size_t* symm_bytes) {
    using namespace stx;
    if (n_local < 0 || m_cap < 0 || world < 1 || world > kMaxWorld || D <= 0 || !ws_bytes || !symm_bytes) {
        set_error("stx_cosine_gather_sizes: bad argument (1 <= world <= %d)", kMaxWorld);
        return STX_EINVAL;
    }
    const int Dp = padded_d(D);
    *ws_bytes = 2 * align256(size_t(n_local) * Dp * sizeof(float)) + 256;
    // two sets of (planes, flags), used alternately by epoch parity, + the push counter
    *symm_bytes = 2 * (symm_planes_bytes(world, m_cap, Dp) + 256) + 256;
    return 0;
}

int stx_cosine_nxm_gathered(const float* d_a, const float* d_b, int n_local, int D, int world, int rank,
                            const int32_t* h_counts, int m_cap, void* const* h_peer_symm, void* d_multicast,
                            uint32_t epoch, float* d_S, void* d_ws, size_t ws_bytes, void* stream) {
    using namespace stx;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || n_local < 0 || D <= 0 || m_cap <= 0 || !h_counts ||
        !h_peer_symm) {
        set_error("stx_cosine_nxm_gathered: bad argument");
        return STX_EINVAL;
    }
    int M = 0;
    for (int r = 0; r < world; ++r) {
        if (h_counts[r] < 0 || h_counts[r] > m_cap || !h_peer_symm[r]) { set_error("stx_cosine_nxm_gathered: bad shard %d", r); return STX_EINVAL; }
        M += h_counts[r];
    }
    if ((n_local > 0 && !d_a) || (h_counts[rank] > 0 && !d_b) || (n_local > 0 && M > 0 && !d_S) || !d_ws) {
        set_error("stx_cosine_nxm_gathered: null pointer");           // (empty shards may come with null pointers)
        return STX_EINVAL;
    }
    size_t need = 0, symm = 0;
    stx_cosine_gather_sizes(n_local, m_cap, world, D, &need, &symm);
    if (ws_bytes < need) { set_error("stx_cosine_nxm_gathered: workspace %zu < %zu bytes", ws_bytes, need); return STX_ENOSPACE; }
    if (int rc = check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the symmetric buffer is laid out for float32 planes of padded_d(D) columns (stx_cosine_gather_sizes); bfloat16 planes of
    // padded_d64(D) columns use the front of every set
    const bool bf = cosine_bf16();
    const int Dp = bf ? padded_d64(D) : padded_d(D);
    const size_t planes = symm_planes_bytes(world, m_cap, padded_d(D));
    const size_t set_bytes = planes + 256;                  // planes, then the flags of this set
    const size_t set_off = size_t(epoch & 1u) * set_bytes;  // calls alternate between two sets: see stx_b200.h
    const size_t slot_bytes = size_t(2) * m_cap * Dp * (bf ? 2 : 4);
    float* a_planes = static_cast<float*>(d_ws);
    SplitDst dst = {};
    dst.a_planes = a_planes;  dst.a_plane_stride = size_t(n_local) * Dp;
    dst.b_plane_stride = size_t(m_cap) * Dp;  dst.n_dst = world;
    for (int p = 0; p < world; ++p) {
        char* base = static_cast<char*>(h_peer_symm[p]) + set_off;
        dst.b_dst[p] = reinterpret_cast<float*>(base + size_t(rank) * slot_bytes);        // my slot in peer p's buffer
        dst.flags[p] = reinterpret_cast<unsigned*>(base + planes);
    }
    if (d_multicast) dst.b_mc = reinterpret_cast<float*>(static_cast<char*>(d_multicast) + set_off + size_t(rank) * slot_bytes);
    char* mine = static_cast<char*>(h_peer_symm[rank]) + set_off;
    dst.flag_index = rank;  dst.epoch = epoch;
    dst.counter = reinterpret_cast<unsigned*>(static_cast<char*>(h_peer_symm[rank]) + 2 * set_bytes);
    // normalise + split + all-gather over NVLink (P2P stores into every peer's slot) in one kernel
    // (an empty shard still launches one CTA: it has nothing to store but it publishes this rank's flag)
    if (bf) STX_LAUNCH(c_split<true>, dim3(std::max(1, (n_local + h_counts[rank] + 7) / 8)), dim3(256), 0, st, d_a, d_b, n_local,
                       h_counts[rank], D, Dp, nullptr, nullptr, nullptr, dst);
    else STX_LAUNCH(c_split<false>, dim3(std::max(1, (n_local + h_counts[rank] + 7) / 8)), dim3(256), 0, st, d_a, d_b, n_local,
                    h_counts[rank], D, Dp, nullptr, nullptr, nullptr, dst);
    const unsigned* my_flags = reinterpret_cast<const unsigned*>(mine + planes);
    bool empty_slot = false;
    for (int r = 0; r < world; ++r) empty_slot |= h_counts[r] == 0;
    if (n_local == 0 || M == 0 || empty_slot)       // flags the GEMM below will not acquire (see c_acquire_flags)
        STX_LAUNCH(c_acquire_flags, dim3(1), dim3(32), 0, st, my_flags, world, epoch);
    if (n_local == 0 || M == 0) return 0;
    TcGeom g = {};
    g.n_rows = n_local;  g.a_plane_rows = n_local;  g.b_plane_rows = m_cap;  g.world = world;  g.ldS = M;  g.bf16 = bf;
    int col = 0;
    for (int r = 0; r < world; ++r) { g.m_count[r] = h_counts[r]; g.col_start[r] = col; col += h_counts[r]; }
    int tiles = 0;
    for (int o = 0; o < world; ++o) {           // own columns first, then ring order: early CTAs never wait
        const int slot = (rank + o) % world;
        g.slot[o] = slot;
        g.tiles_start[o] = tiles;
        tiles += (h_counts[slot] + kTN - 1) / kTN;
    }
    for (int o = world; o <= kMaxWorld; ++o) g.tiles_start[o] = tiles;
    g.flags = reinterpret_cast<const unsigned*>(mine + planes);
    g.epoch = epoch;
    return launch_gemm(a_planes, 2 * n_local, reinterpret_cast<const float*>(mine), world * 2 * m_cap, Dp, g, tiles, d_S, st);
}

}  // extern "C"
