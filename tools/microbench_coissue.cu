// What does an instruction of each class cost its SCHEDULER?  8 warps per SM: warps 0..3 (one per scheduler) run stream A,
// warps 4..7 (one per scheduler) run stream B; each stream alone, then both together.  If the co-run takes max(a, b) the two
// classes overlap freely; if it takes ~a + b they serialise in the scheduler's dispatch port (or share a pipe).
// Streams: DFMA, FFMA, IADD (integer ALU), F2F f32->f64, F2F f64->f32, LDS.32, LDS.128, STS.128 -- 8 independent chains each.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_coissue tools/microbench_coissue.cu
#include <cuda_runtime.h>
#include <cstdio>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
constexpr int ITERS = 4096;

enum { DFMA, FFMA, IADD, CVT_F2D, CVT_D2F, LDS32, LDS128, STS128, NKIND };
__device__ __forceinline__ double stream(int kind, float4* buf, int seed) {
    double res = 0;
    switch (kind) {
    case DFMA: {
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = seed + i;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fma(v[i], 1.0001, 0.5);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) res += v[i];
    } break;
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
    case IADD: {
        unsigned v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = seed + i;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = v[i] * 3u + 0x9e3779b9u;       // one IMAD
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) res += v[i];
    } break;
    case CVT_F2D: {
        float v[8]; double acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] = seed + i; acc[i] = 0; }
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double d = (double)v[i];
                acc[i] = __longlong_as_double(__double_as_longlong(acc[i]) ^ __double_as_longlong(d));   // 2 LOP3
                v[i] = __int_as_float(__float_as_int(v[i]) + 1);                                          // 1 IADD
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) res += acc[i];
    } break;
    case CVT_D2F: {
        double v[8]; int acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] = seed + i; acc[i] = 0; }
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i] ^= __float_as_int((float)v[i]);
                v[i] = __longlong_as_double(__double_as_longlong(v[i]) + 1);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) res += acc[i];
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
    case LDS128: {
        float acc = 0; int idx = seed & 31;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += buf[(idx + 32 * i) & 511].x;
            idx += 7;
        }
        res = acc;
    } break;
    case STS128: {
        int idx = seed & 31;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) buf[512 + ((idx + 32 * i) & 511)] = make_float4(it, i, 0.f, 0.f);
            idx += 7;
        }
        res = buf[512 + (seed & 511)].x;
    } break;
    }
    return res;
}

__global__ void __launch_bounds__(256) k_pair(double* out, int ka, int kb) {
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, 0.f, 0.f, 0.f);
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
    const char* names[NKIND] = {"DFMA", "FFMA", "IMAD", "F2F f32->f64 (+3 ALU)", "F2F f64->f32 (+2 ALU)", "LDS.32 (+FADD)", "LDS.128 (+FADD)", "STS.128"};
    float alone[NKIND];
    printf("one warp per scheduler, 8 chains, %d instructions of the class per warp; cycles per instruction of the class:\n", ITERS * 8);
    for (int k = 0; k < NKIND; ++k) {
        alone[k] = time_ms([&] { k_pair<<<sms, 256>>>(out, k, -1); });
        printf("  %-24s alone %.3f ms = %.2f cycles per instruction\n", names[k], alone[k], alone[k] * 1e-3 * clk_khz * 1e3 / (ITERS * 8.0));
    }
    printf("pairs on the same scheduler: together vs max and sum of the two alone\n");
    for (int a : {DFMA, FFMA})
        for (int b = 0; b < NKIND; ++b) {
            if (b == a) continue;
            const float t = time_ms([&] { k_pair<<<sms, 256>>>(out, a, b); });
            const float mx = alone[a] > alone[b] ? alone[a] : alone[b], sm = alone[a] + alone[b];
            printf("  %-5s + %-24s together %.3f ms   max %.3f  sum %.3f   overlap %.0f %%\n", names[a], names[b], t, mx, sm,
                   100.0 * (sm - t) / (sm - mx > 1e-6 ? sm - mx : 1e-6));
        }
    CHECK(cudaDeviceSynchronize()); CHECK(cudaGetLastError());
    return 0;
}
