// Can two half-CTA warp groups in different phases overlap the FP64 pipe and the shared-memory pipe on one SM?
// Synthetic FFT-core of recipe K: phase F1 = window + k_pass1 codelet + twiddles (FP64), S = 16 x STS.128 into the exchange,
// L = 16 x LDS.128 from it, F2 = dft16 codelet + power (FP64).  Same work per 32-frame tile in both modes:
//   mode 0  16 warps x 1 role, CTA barriers (the shipped k_frames structure)
//   mode 1  2 groups x 8 warps x 2 roles, one tile per group, group barriers, group 1 shifted by half a period
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_overlap tools/microbench_overlap.cu
#include <cuda_runtime.h>
#include <cstdio>
#include "../speech_transcript_embeddings_b200/csrc/codelets.cuh"
using namespace stx;

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__constant__ double c_win[16][25];
__constant__ double2 c_tw[16][16];

__device__ __forceinline__ void group_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int SLOTS>
__device__ __forceinline__ void f1_store(double2 (*ex)[SLOTS][32], int role, int slot, int lane, double& seed) {
    double y[25], re[17], im[17];
#pragma unroll
    for (int i = 0; i < 25; ++i) y[i] = c_win[role][i] * (seed + i);
    codelets::k_pass1<double>(y, re, im);
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) {
        const double2 t = c_tw[role][k1];
        ex[k1][slot][lane] = make_double2(fma(re[k1], t.x, -(im[k1] * t.y)), fma(re[k1], t.y, im[k1] * t.x));
    }
    ex[0][slot][lane] = make_double2(re[0], re[16]);
    seed += re[3];
}
template <int SLOTS>
__device__ __forceinline__ void load_f2(double2 (*ex)[SLOTS][32], int row, int lane, double& acc) {
    double xr[16], xi[16], yr[16], yi[16];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) { const double2 v = ex[row][n2 % SLOTS][lane]; xr[n2] = v.x; xi[n2] = v.y; }
    codelets::dft16<double>(xr, xi, yr, yi);
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += fma(yr[k], yr[k], yi[k] * yi[k]);
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_overlap(double* out, int tiles) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    double seed = 1.0 + 1e-3 * threadIdx.x, acc = 0.0;
    if (MODE == 0) {
        double2 (*ex)[16][32] = reinterpret_cast<double2 (*)[16][32]>(raw);       // [16][16][32] = 128 KB
        for (int t = 0; t < tiles; ++t) {
            f1_store<16>(ex, warp, warp, lane, seed);
            __syncthreads();
            load_f2<16>(ex, warp, lane, acc);
            __syncthreads();
        }
    } else {
        const int g = warp >> 3, w8 = warp & 7;
        double2 (*ex)[8][32] = reinterpret_cast<double2 (*)[8][32]>(raw + g * 65536);   // per group [16][8][32] = 64 KB: both roles of a
                                                                                       // warp reuse slot w8 (only the traffic matters here)
        if (g == 1) f1_store<8>(ex, w8, w8, lane, seed);      // half a period late: its FP64 phase meets group 0's memory phase
        for (int t = 0; t < tiles / 2; ++t) {
            f1_store<8>(ex, w8, w8, lane, seed);
            f1_store<8>(ex, w8 + 8, w8, lane, seed);
            group_bar(1 + g, 256);
            load_f2<8>(ex, w8, lane, acc);
            load_f2<8>(ex, w8 + 8, lane, acc);
            group_bar(1 + g, 256);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + seed;
}

template <typename F> float time_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double win[16][25]; double2 tw[16][16];
    for (int i = 0; i < 16; ++i) { for (int j = 0; j < 25; ++j) win[i][j] = 0.5 + 0.01 * (i + j); for (int j = 0; j < 16; ++j) tw[i][j] = make_double2(0.6, 0.8); }
    CHECK(cudaMemcpyToSymbol(c_win, win, sizeof(win))); CHECK(cudaMemcpyToSymbol(c_tw, tw, sizeof(tw)));
    double* out; CHECK(cudaMalloc(&out, size_t(p.multiProcessorCount) * 512 * 8));
    CHECK(cudaFuncSetAttribute(k_overlap<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    CHECK(cudaFuncSetAttribute(k_overlap<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    const int tiles = 4096;
    const float ms0 = time_ms([&] { k_overlap<0><<<p.multiProcessorCount, 512, 131072>>>(out, tiles); });
    const float ms1 = time_ms([&] { k_overlap<1><<<p.multiProcessorCount, 512, 131072>>>(out, tiles); });
    CHECK(cudaGetLastError());
    printf("FFT-core of a 32-frame tile, cycles per tile per SM @%.0f MHz:  lockstep (16 warps x 1 role) %.0f   two staggered groups (8 warps x 2 roles) %.0f   ratio %.2f\n",
           clk_khz / 1e3, ms0 * 1e-3 * clk_khz * 1e3 / tiles, ms1 * 1e-3 * clk_khz * 1e3 / tiles, ms0 / ms1);
    return 0;
}
