// How many warps and how much instruction-level parallelism does the FP64 pipe of one SM need?
//   (1) DFMA chains: lane-ops per clock per SM for 1 / 2 / 4 warps per scheduler and 1 / 2 / 4 / 8 independent chains per thread
//       (the latency of a dependent DFMA follows from the ILP = 1 row)
//   (2) the generated codelets (k_pass1, dft16) on register data with 8 and 16 warps per SM at the 128-register budget
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_fp64 tools/microbench_fp64.cu
#include <cuda_runtime.h>
#include <cstdio>
#include "../speech_transcript_embeddings_b200/csrc/codelets.cuh"

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
constexpr int ITERS = 8192;

template <int ILP>
__global__ void k_dfma(double* out, double a, double b) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int kWhich>
__global__ void __launch_bounds__(512, 1) k_codelet(double* out, double seed, int iters) {
    using namespace stx::codelets;
    double y[25], re[17], im[17];
#pragma unroll
    for (int i = 0; i < 25; ++i) y[i] = seed * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 17; ++i) re[i] = im[i] = 0.0;
    double xr[16], xi[16], yr[16], yi[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { xr[i] = seed * i; xi[i] = seed + i; }
    for (int it = 0; it < iters; ++it) {
        if (kWhich == 0) {
            k_pass1<double>(y, re, im);
#pragma unroll
            for (int i = 0; i < 25; ++i) y[i] = re[i % 17] * 0.5 + (i < 15 ? im[1 + i] : y[i]) * 0.25;
        } else {
            dft16<double>(xr, xi, yr, yi);
#pragma unroll
            for (int i = 0; i < 16; ++i) { xr[i] = yr[i] * 0.5; xi[i] = yi[i] * 0.5; }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 25; ++i) s += y[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += xr[i] + xi[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
    double* out; CHECK(cudaMalloc(&out, size_t(sms) * 1024 * 8));
    printf("DFMA lane-ops/clk/SM (peak 64) and cycles per dependent DFMA of one warp; %d SMs @%.0f MHz\n", sms, clk_khz / 1e3);
    for (int threads : {128, 256, 512}) {
        auto rep = [&](int ilp, float ms) {
            const double ops = double(sms) * threads * ITERS * ilp;
            const double per_clk = ops / (ms * 1e-3) / sms / (clk_khz * 1e3);
            printf("  %2d warps/SM  ILP %d: %6.1f lane-ops/clk/SM   (%.1f cycles per DFMA per warp)\n", threads / 32, ilp, per_clk,
                   ms * 1e-3 * clk_khz * 1e3 / (double(ITERS) * ilp));
        };
        rep(1, time_ms([&] { k_dfma<1><<<sms, threads>>>(out, 1.0001, 0.5); }));
        rep(2, time_ms([&] { k_dfma<2><<<sms, threads>>>(out, 1.0001, 0.5); }));
        rep(4, time_ms([&] { k_dfma<4><<<sms, threads>>>(out, 1.0001, 0.5); }));
        rep(8, time_ms([&] { k_dfma<8><<<sms, threads>>>(out, 1.0001, 0.5); }));
    }
    const int it = 2000;
    for (int threads : {128, 256, 512}) {
        auto rep2 = [&](const char* name, float ms, double ops) {
            const double per_clk = double(sms) * threads * it * ops / (ms * 1e-3) / sms / (clk_khz * 1e3);
            printf("  %-10s %2d warps/SM: %6.1f lane-ops/clk/SM  (%.0f cycles per codelet call per warp)\n", name, threads / 32, per_clk,
                   ms * 1e-3 * clk_khz * 1e3 / it);
        };
        rep2("k_pass1", time_ms([&] { k_codelet<0><<<sms, threads>>>(out, 1e-3, it); }), 144 + 50);
        rep2("dft16", time_ms([&] { k_codelet<1><<<sms, threads>>>(out, 1e-3, it); }), 144 + 32);
    }
    CHECK(cudaDeviceSynchronize()); CHECK(cudaGetLastError());
    return 0;
}
