// Staging strategies for pageable -> device: does a small ring of pinned slots that stays in the last-level cache (regular stores,
// the DMA engine reading the lines back from cache) beat one batch-sized pinned buffer written with non-temporal stores?
//   A  one 123 MB pinned buffer, non-temporal stores, chunk by chunk, cudaMemcpyAsync per chunk (what the library does)
//   B  ring of R pinned slots of one chunk each, REGULAR stores; a slot is reused once its H2D copy has completed
//   C  ring, non-temporal stores
//   D  one 123 MB WRITE-COMBINED pinned buffer (cudaHostAllocWriteCombined: not snooped by PCIe reads), non-temporal stores
//   E  the H2D copies alone (buffer already packed)        F  the packing alone (no copies)
// 64 pageable source buffers of 30 s (1.92 MB each), T packing threads, 62 MB of D2H traffic in flight on another stream.
// Build: nvcc -O3 -std=c++17 -o tools/microbench_staging tools/microbench_staging.cu -Xcompiler -mavx2,-pthread
#include <cuda_runtime.h>
#include <immintrin.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static void copy_nt(unsigned char* dst, const unsigned char* src, size_t n) {
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
        _mm256_stream_si256((__m256i*)(dst + i), a); _mm256_stream_si256((__m256i*)(dst + i + 32), b);
        _mm256_stream_si256((__m256i*)(dst + i + 64), c); _mm256_stream_si256((__m256i*)(dst + i + 96), d);
    }
    memcpy(dst + i, src + i, n - i);
    _mm_sfence();
}
static void copy_reg(unsigned char* dst, const unsigned char* src, size_t n) { memcpy(dst, src, n); }

struct Pool {       // fork-join over T threads (spawned per call: ~20 us, negligible against the 0.1-0.3 ms a chunk takes)
    int T;
    template <typename F> void run(F f) {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back([&, t] { f(t); });
        f(0);
        for (auto& x : th) x.join();
    }
};

int main(int argc, char** argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 8;
    const size_t clip = 480000 * 4, B = 64, total = clip * B;
    const size_t chunk = argc > 2 ? (size_t)atoi(argv[2]) * clip : 3 * clip;       // clips per chunk
    const int R = argc > 3 ? atoi(argv[3]) : 4;
    std::vector<unsigned char*> src(B);
    for (auto& p : src) { p = (unsigned char*)malloc(clip); memset(p, 1, clip); }
    unsigned char *big, *bigwc, *ring, *hout, *dpcm, *dout;
    CHECK(cudaHostAlloc(&bigwc, total, cudaHostAllocWriteCombined));
    CHECK(cudaMallocHost(&big, total)); CHECK(cudaMallocHost(&ring, chunk * R)); CHECK(cudaMallocHost(&hout, total / 2));
    CHECK(cudaMalloc(&dpcm, total)); CHECK(cudaMalloc(&dout, total / 2));
    memset(big, 0, total); memset(ring, 0, chunk * R); memset(hout, 0, total / 2);
    cudaStream_t s_in, s_out; CHECK(cudaStreamCreate(&s_in)); CHECK(cudaStreamCreate(&s_out));
    const int nchunks = (int)((total + chunk - 1) / chunk);
    std::vector<cudaEvent_t> ev(nchunks);
    for (auto& evt : ev) CHECK(cudaEventCreateWithFlags(&evt, cudaEventDisableTiming));
    Pool pool{T};
    auto pack_chunk = [&](unsigned char* dst, size_t lo, size_t hi, bool nt) {     // bytes [lo, hi) of the concatenated clips
        pool.run([&](int t) {
            const size_t n = hi - lo, per = ((n + T - 1) / T + 63) & ~size_t(63);
            const size_t a = std::min(n, per * t), b = std::min(n, per * (t + 1));
            size_t pos = lo + a;
            while (pos < lo + b) {
                const size_t c = pos / clip, off = pos % clip, len = std::min(clip - off, lo + b - pos);
                (nt ? copy_nt : copy_reg)(dst + (pos - lo), src[c] + off, len);
                pos += len;
            }
        });
    };
    auto run = [&](int mode) {
        // D2H traffic of the previous step's features, in 8 pieces, concurrently
        for (int i = 0; i < 8; ++i) CHECK(cudaMemcpyAsync(hout + i * (total / 16), dout + i * (total / 16), total / 16, cudaMemcpyDeviceToHost, s_out));
        for (int c = 0; c < nchunks; ++c) {
            const size_t lo = c * chunk, hi = std::min(total, lo + chunk);
            if (mode == 0 || mode >= 3) {
                unsigned char* buf = mode == 3 ? bigwc : big;
                if (mode != 4) pack_chunk(buf + lo, lo, hi, true);
                if (mode != 5) CHECK(cudaMemcpyAsync(dpcm + lo, buf + lo, hi - lo, cudaMemcpyHostToDevice, s_in));
            } else {
                if (c >= R) CHECK(cudaEventSynchronize(ev[c - R]));
                unsigned char* slot = ring + (c % R) * chunk;
                pack_chunk(slot, lo, hi, mode == 2);
                CHECK(cudaMemcpyAsync(dpcm + lo, slot, hi - lo, cudaMemcpyHostToDevice, s_in));
                CHECK(cudaEventRecord(ev[c], s_in));
            }
        }
        CHECK(cudaStreamSynchronize(s_in)); CHECK(cudaStreamSynchronize(s_out));
    };
    const char* names[6] = {"A  one big buffer, non-temporal stores", "B  ring, regular stores", "C  ring, non-temporal stores",
                            "D  one big WRITE-COMBINED buffer, NT stores", "E  H2D copies alone", "F  packing alone"};
    printf("%d threads, chunk %.1f MB, ring of %d slots (%.1f MB)\n", T, chunk / 1e6, R, chunk * R / 1e6);
    for (int rep = 0; rep < 2; ++rep)
        for (int mode = 0; mode < 6; ++mode) {
            run(mode); run(mode);
            const auto t0 = std::chrono::steady_clock::now();
            const int iters = 10;
            for (int i = 0; i < iters; ++i) run(mode);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / iters;
            printf("  %-42s %.3f ms per 123 MB batch (%.1f GB/s into the device)\n", names[mode], ms, total / ms / 1e6);
        }
    return 0;
}
