// Packed float32 arithmetic on sm_100a: what does an FFMA2 (fma.rn.f32x2: two float32 FMAs per lane and instruction) cost its
// SCHEDULER, and does it leave the dispatch port free for other classes the way a scalar FFMA does?  Same frame as
// tools/microbench_coissue.cu: 8 warps per SM, warps 0..3 (one per scheduler) run stream A, warps 4..7 stream B; each stream
// alone, then both together.  Streams: FFMA, FFMA2, FADD2, FMUL2, IMAD, LDS.32, LDS.64 (8 independent chains each), and the
// two-warp forms FFMA+FFMA / FFMA2+FFMA2 (two warps of the same class per scheduler).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_ffma2 tools/microbench_ffma2.cu
#include <cuda_runtime.h>
#include <cstdio>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
constexpr int ITERS = 4096;

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    return ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a);
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

enum { FFMA, FFMA2, FADD2, FMUL2, IMAD, LDS32, LDS64, NKIND };
__device__ __forceinline__ double stream(int kind, float2* buf, int seed) {
    double res = 0;
    switch (kind) {
    case FFMA: {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = seed + i;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], 1.0001f, 0.5f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) res += v[i];
    } break;
    case FFMA2: case FADD2: case FMUL2: {
        unsigned long long v[8];
        const unsigned long long m = pk(1.0001f, 0.9999f), c = pk(0.5f, 0.25f);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = pk(seed + i, seed - i);
        if (kind == FFMA2)
            for (int it = 0; it < ITERS; ++it) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fma2(v[i], m, c);
            }
        else if (kind == FADD2)
            for (int it = 0; it < ITERS; ++it) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = add2(v[i], c);
            }
        else
            for (int it = 0; it < ITERS; ++it) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = mul2(v[i], m);
            }
#pragma unroll
        for (int i = 0; i < 8; ++i) res += __uint_as_float((unsigned)v[i]) + __uint_as_float((unsigned)(v[i] >> 32));
    } break;
    case IMAD: {
        unsigned v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = seed + i;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = v[i] * 3u + 0x9e3779b9u;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) res += v[i];
    } break;
    case LDS32: {
        const float* b = reinterpret_cast<const float*>(buf);
        float acc = 0; int idx = seed & 31;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += b[(idx + 32 * i) & 2047];
            idx += 7;
        }
        res = acc;
    } break;
    case LDS64: {
        float acc = 0; int idx = seed & 31;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float2 t = buf[(idx + 32 * i) & 1023]; acc += t.x; }
            idx += 7;
        }
        res = acc;
    } break;
    }
    return res;
}

__global__ void __launch_bounds__(256) k_pair(double* out, int ka, int kb) {
    __shared__ float2 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_float2(i, 0.f);
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int kind = (warp & 4) == 0 ? ka : kb;        // scheduler = warp % 4: one A warp and one B warp per scheduler
    double r = 0;
    if (kind >= 0) r = stream(kind, buf, threadIdx.x);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F> static float time_ms(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch();
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    double* out; CHECK(cudaMalloc(&out, size_t(sms) * 256 * 8));
    const char* names[NKIND] = {"FFMA", "FFMA2", "FADD2", "FMUL2", "IMAD", "LDS.32 (+FADD)", "LDS.64 (+FADD)"};
    float alone[NKIND];
    printf("one warp per scheduler, 8 chains, %d instructions of the class per warp; cycles per instruction of the class:\n", ITERS * 8);
    for (int k = 0; k < NKIND; ++k) {
        alone[k] = time_ms([&] { k_pair<<<sms, 256>>>(out, k, -1); });
        printf("  %-16s alone %.3f ms = %.2f cycles per instruction\n", names[k], alone[k], alone[k] * 1e-3 * clk_khz * 1e3 / (ITERS * 8.0));
    }
    printf("two warps per scheduler, together vs max and sum of the two alone (overlap 100 %% = free co-issue, 0 %% = serialised)\n");
    for (int a : {FFMA, FFMA2, FADD2})
        for (int b = 0; b < NKIND; ++b) {
            const float t = time_ms([&] { k_pair<<<sms, 256>>>(out, a, b); });
            const float mx = alone[a] > alone[b] ? alone[a] : alone[b], sm = alone[a] + alone[b];
            printf("  %-5s + %-16s together %.3f ms   max %.3f  sum %.3f   overlap %.0f %%\n", names[a], names[b], t, mx, sm,
                   100.0 * (sm - t) / (sm - mx > 1e-6 ? sm - mx : 1e-6));
        }
    CHECK(cudaDeviceSynchronize()); CHECK(cudaGetLastError());
    return 0;
}
