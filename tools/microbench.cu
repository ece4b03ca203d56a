// Pipe-throughput microbenchmarks used to size the kernels (FP64 vs FP32 pipes, conversions, shuffles,
// shared memory).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Prints lane-operations per clock per SM for each instruction class.
#include <cuda_runtime.h>
#include <cstdio>
#include "../speech_transcript_embeddings_b200/csrc/codelets.cuh"

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <typename T>
__global__ void k_fma(T* out, T a, T b) {
    T v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = v[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename T>
__global__ void k_add(T* out, T a) {
    T v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = v[i] + a;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_cvt_f2d(double* out, float a) {
    float v[ILP];
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { v[i] = threadIdx.x + i + a; acc[i] = 0; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            double d = (double)v[i];                       // F2F.F64.F32
            acc[i] = __longlong_as_double(__double_as_longlong(acc[i]) ^ __double_as_longlong(d));   // cheap integer consumer
            v[i] = __int_as_float(__float_as_int(v[i]) + 1);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_cvt_d2f(float* out, double a) {
    double v[ILP];
    int acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { v[i] = threadIdx.x + i + a; acc[i] = 0; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float f = (float)v[i];                         // F2F.F32.F64
            acc[i] ^= __float_as_int(f);
            v[i] = __longlong_as_double(__double_as_longlong(v[i]) + 1);
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __int_as_float(s);
}

__global__ void k_shfl(int* out) {
    int v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1 + (i & 7));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename V>
__global__ void k_lds(float* out) {
    __shared__ V buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = V{};
    __syncthreads();
    float acc = 0;
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            V x = buf[(idx + i * 32) & 1023];
            acc += *reinterpret_cast<float*>(&x);
        }
        idx += 7;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// shared loads with duplicated addresses: how many wavefronts does a table read cost when both
// half-warps (or neighbouring lanes) read the same elements?  mode 0: idx = lane, 1: lane & 15, 2: lane >> 1
template <typename V, int kMode>
__global__ void k_lds_dup(float* out) {
    __shared__ V buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = V{};
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int base = kMode == 0 ? lane : (kMode == 1 ? (lane & 15) : (lane >> 1));
    float acc = 0;
    int idx = base;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            V x = buf[(idx + i * 32) & 2047];
            acc += *reinterpret_cast<float*>(&x);
        }
        idx += 32 * ILP;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch();
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// The generated codelets on register data only (no memory): how close does the compiler's schedule of the
// straight-line FP64 code get to the pipe's peak with 16 warps per SM and <= 128 registers (the k_frames budget)?
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
            for (int i = 0; i < 25; ++i) y[i] = re[i % 17] * 0.5 + (i < 15 ? im[1 + i] : y[i]) * 0.25;   // +50 ops, keeps everything live
        } else {
            dft16<double>(xr, xi, yr, yi);
#pragma unroll
            for (int i = 0; i < 16; ++i) { xr[i] = yr[i] * 0.5; xi[i] = yi[i] * 0.5; }                  // +32 ops
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 25; ++i) s += y[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += xr[i] + xi[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s sm_%d%d, %d SMs, clock attr %.0f MHz\n", p.name, p.major, p.minor, p.multiProcessorCount, clk_khz / 1e3);
    const int blocks = p.multiProcessorCount * 8, threads = 256;
    void* out;
    CHECK(cudaMalloc(&out, size_t(blocks) * threads * 8));
    const double lanes = double(blocks) * threads * ITERS * ILP;
    auto report = [&](const char* name, float ms, double per_lane_ops) {
        double ops_per_s = lanes * per_lane_ops / (ms * 1e-3);
        printf("%-22s %8.3f ms  %8.2f Tops/s  %7.1f lane-ops/clk/SM @%.0f MHz (attr clock)\n", name, ms, ops_per_s / 1e12,
               ops_per_s / p.multiProcessorCount / (clk_khz * 1e3), clk_khz / 1e3);
    };
    report("FFMA", time_ms([&] { k_fma<float><<<blocks, threads>>>((float*)out, 1.0001f, 0.5f); }), 1);
    report("FADD", time_ms([&] { k_add<float><<<blocks, threads>>>((float*)out, 0.5f); }), 1);
    report("DFMA", time_ms([&] { k_fma<double><<<blocks, threads>>>((double*)out, 1.0001, 0.5); }), 1);
    report("DADD", time_ms([&] { k_add<double><<<blocks, threads>>>((double*)out, 0.5); }), 1);
    report("F2F f32->f64", time_ms([&] { k_cvt_f2d<<<blocks, threads>>>((double*)out, 0.5f); }), 1);
    report("F2F f64->f32", time_ms([&] { k_cvt_d2f<<<blocks, threads>>>((float*)out, 0.5); }), 1);
    report("SHFL.32", time_ms([&] { k_shfl<<<blocks, threads>>>((int*)out); }), 1);
    report("LDS.32", time_ms([&] { k_lds<float><<<blocks, threads>>>((float*)out); }), 1);
    report("LDS.64", time_ms([&] { k_lds<float2><<<blocks, threads>>>((float*)out); }), 1);
    report("LDS.128", time_ms([&] { k_lds<float4><<<blocks, threads>>>((float*)out); }), 1);
    report("LDS.128 idx=lane", time_ms([&] { k_lds_dup<float4, 0><<<blocks, threads>>>((float*)out); }), 1);
    report("LDS.128 idx=lane&15", time_ms([&] { k_lds_dup<float4, 1><<<blocks, threads>>>((float*)out); }), 1);
    report("LDS.128 idx=lane>>1", time_ms([&] { k_lds_dup<float4, 2><<<blocks, threads>>>((float*)out); }), 1);
    report("LDS.64 idx=lane", time_ms([&] { k_lds_dup<float2, 0><<<blocks, threads>>>((float*)out); }), 1);
    report("LDS.64 idx=lane&15", time_ms([&] { k_lds_dup<float2, 1><<<blocks, threads>>>((float*)out); }), 1);
    report("LDS.64 idx=lane>>1", time_ms([&] { k_lds_dup<float2, 2><<<blocks, threads>>>((float*)out); }), 1);
    report("LDS.32 idx=lane&15", time_ms([&] { k_lds_dup<float, 1><<<blocks, threads>>>((float*)out); }), 1);
    {
        // ops per iteration from tools/gen_codelets.py (after FMA contraction) + the feedback ops above
        const int it = 2000, cb = p.multiProcessorCount, ct = 512;
        void* o2;
        CHECK(cudaMalloc(&o2, size_t(cb) * ct * 8));
        const double l2 = double(cb) * ct * it;
        auto rep2 = [&](const char* name, float ms, double ops) {
            double ops_per_s = l2 * ops / (ms * 1e-3);
            printf("%-22s %8.3f ms  %8.2f Tops/s  %7.1f lane-ops/clk/SM @%.0f MHz (attr clock)\n", name, ms, ops_per_s / 1e12,
                   ops_per_s / p.multiProcessorCount / (clk_khz * 1e3), clk_khz / 1e3);
        };
        rep2("codelet k_pass1 f64", time_ms([&] { k_codelet<0><<<cb, ct>>>((double*)o2, 1e-3, it); }), 225 + 50);
        rep2("codelet dft16 f64", time_ms([&] { k_codelet<1><<<cb, ct>>>((double*)o2, 1e-3, it); }), 176 + 32);
    }
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaGetLastError());
    return 0;
}
