// Does a warp that saturates the FP64 pipe leave issue slots for the other warps of its scheduler?
// 8 warps per SM (2 per scheduler): warps 0..3 run a DFMA stream (ILP 8), warps 4..7 run an FP32 / integer / shared-memory
// stream.  Each stream alone, then both together: if the co-run takes max(a, b) the pipes overlap freely; if it takes
// ~a + b the FP64 instructions hold the scheduler's dispatch port.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_coissue tools/microbench_coissue.cu
#include <cuda_runtime.h>
#include <cstdio>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
constexpr int ITERS = 8192;

// kind: 0 FFMA, 1 IADD3-ish integer, 2 LDS.32
template <int KIND>
__global__ void __launch_bounds__(256) k_mix(double* out, int do_f64, int do_other, double a, double b) {
    __shared__ float buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = (float)i;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    double res = 0;
    if ((warp & 4) == 0) {      // warps 0..3: one per scheduler (scheduler = warp % 4); warps 4..7 run the other stream
        if (do_f64) {
            double v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
            for (int it = 0; it < ITERS; ++it) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fma(v[i], a, b);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) res += v[i];
        }
    } else if (do_other) {
        if (KIND == 0) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
            const float fa = (float)a, fb = (float)b;
            for (int it = 0; it < 2 * ITERS; ++it) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], fa, fb);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) res += v[i];
        } else if (KIND == 1) {
            unsigned v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
            for (int it = 0; it < 2 * ITERS; ++it) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (v[i] ^ (unsigned)it) + 0x9e3779b9u;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) res += v[i];
        } else {
            float acc = 0;
            int idx = threadIdx.x & 31;
            for (int it = 0; it < ITERS; ++it) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc += buf[(idx + 32 * i) & 2047];
                idx += 7;
            }
            res = acc;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
}

template <typename F> static float time_ms(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch();
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* out; CHECK(cudaMalloc(&out, size_t(sms) * 256 * 8));
    const char* names[3] = {"FFMA x2", "integer x2 (2 ops)", "LDS.32 + FADD"};
    for (int kind = 0; kind < 3; ++kind) {
        auto run = [&](int f, int o) {
            return time_ms([&] {
                if (kind == 0) k_mix<0><<<sms, 256>>>(out, f, o, 1.0001, 0.5);
                else if (kind == 1) k_mix<1><<<sms, 256>>>(out, f, o, 1.0001, 0.5);
                else k_mix<2><<<sms, 256>>>(out, f, o, 1.0001, 0.5);
            });
        };
        const float a = run(1, 0), b = run(0, 1), c = run(1, 1);
        printf("DFMA stream alone %.3f ms | %-20s alone %.3f ms | together %.3f ms  (max %.3f, sum %.3f)\n", a, names[kind], b, c,
               a > b ? a : b, a + b);
    }
    CHECK(cudaDeviceSynchronize()); CHECK(cudaGetLastError());
    return 0;
}
