// Shared host-side plumbing of libstx_b200: error reporting, launch counting, float64 tables.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/stx_b200.h"

namespace stx {

void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);
extern std::atomic<uint64_t> g_launches;

#define STX_CUDA(call)                                             \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return stx::cuda_fail(e__, #call); \
    } while (0)

// Optional per-launch timing (stx_profile_enable): CUDA events recorded on the launching stream
// immediately before and after each kernel; collected by stx_profile_collect.
extern std::atomic<int> g_profile;
void profile_before(const char* name, cudaStream_t st);
void profile_after(cudaStream_t st);

// every kernel launch of the library goes through this macro so that the launch counter is exact
#define STX_LAUNCH(kernel, grid, block, smem, stream, ...)                       \
    do {                                                                         \
        const bool prof__ = stx::g_profile.load(std::memory_order_relaxed) != 0; \
        if (prof__) stx::profile_before(#kernel, (stream));                      \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);              \
        if (prof__) stx::profile_after((stream));                                \
        stx::g_launches.fetch_add(1, std::memory_order_relaxed);                 \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) return stx::cuda_fail(e__, "launch " #kernel);   \
    } while (0)

// ---- float64 host tables (regenerated here, not read from anywhere) -------------------------
const std::vector<double>& k_window();   // [400]  Povey
const std::vector<double>& k_mel();      // [257*80] Kaldi mel, row-major [bin][mel]
const std::vector<double>& w_window();   // [400]  periodic Hann
const std::vector<double>& w_mel();      // [201*80] Slaney mel, row-major [bin][mel]

// Compressed rows of a [nbins, nmel] filterbank: for mel m the non-zero weights are the
// contiguous bins [first[m], first[m]+count[m]) stored at weights[offset[m]...].
struct MelCsr {
    std::vector<int>   first, count, offset;
    std::vector<float> weights;
};
// `pad` > 1 pads every row to a multiple of `pad` weights (zeros) starting at a multiple of `pad` in
// `weights`, and keeps first + count <= limit so that padded entries still index valid bins.
MelCsr build_mel_csr(const std::vector<double>& fb, int nbins, int nmel, double scale, int pad = 1, int limit = 1 << 30);

int check_device();   // 0 if the current device is sm_100, else STX_EDEVICE

// cosine.cu: Linear over rows whose hi / lo TF32 planes are already written (see there)
int project_from_planes(const float* a_planes, int rows, int in_dim, const float* d_weight, const float* d_bias, int out_dim,
                        float* b_planes, float* d_hidden, cudaStream_t st);


}  // namespace stx
