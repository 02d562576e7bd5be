// Hand-off copy probe: how fast does the int16 result of one receiver ([1024][240000], 179968 demodulated columns)
// reach pinned host memory?  (a) one 1-D cudaMemcpyAsync of the same byte count, (b) the 2-D copy the library issues
// (row = 359 936 B, pitch 480 000 B), (c) the 2-D copy split over 2 / 4 streams, (d) a kernel that stores straight
// into the mapped pinned buffer (zero-copy, 16-byte stores, one CTA per row segment).
// (e) the 2-D copy with page-aligned host / device pitches.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/d2h_probe tools/d2h_probe.cu ; prints one JSON line.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));             \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

__global__ void store_rows(const int4* __restrict__ src, int4* __restrict__ dst, size_t pitch16, size_t cols16, int rows) {
    // grid-stride over 16-byte words of the dirty columns of every row; consecutive threads -> consecutive words
    const size_t total = (size_t)rows * cols16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / cols16, c = i - r * cols16;
        dst[r * pitch16 + c] = src[r * pitch16 + c];
    }
}

int main() {
    const size_t rows = 1024, af = 240000, cols = 179968;
    const size_t pitch = af * 2, width = cols * 2, bytes = rows * pitch;
    const int reps = 8, nbuf = 4;
    void* d = nullptr;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMemset(d, 1, bytes));
    std::vector<void*> h(nbuf);
    for (auto& p : h) CK(cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
    cudaStream_t st[4];
    for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    auto timed = [&](auto&& body) {
        body();
        cudaDeviceSynchronize();
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; ++r) body();
        cudaDeviceSynchronize();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return (double)reps * rows * width / s / 1e9;
    };
    int k = 0;
    const double g1d = timed([&] { cudaMemcpyAsync(h[k++ % nbuf], d, rows * width, cudaMemcpyDeviceToHost, st[0]); });
    const double g2d = timed([&] { cudaMemcpy2DAsync(h[k++ % nbuf], pitch, d, pitch, width, rows, cudaMemcpyDeviceToHost, st[0]); });
    auto split = [&](int n) {
        return timed([&] {
            char* hb = (char*)h[k++ % nbuf];
            for (int j = 0; j < n; ++j)
                cudaMemcpy2DAsync(hb + (size_t)j * (rows / n) * pitch, pitch, (char*)d + (size_t)j * (rows / n) * pitch, pitch,
                                  width, rows / n, cudaMemcpyDeviceToHost, st[j]);
        });
    };
    const double g2d2 = split(2), g2d4 = split(4);
    // four whole receivers in flight on four streams (what bench.py's e2e leg does)
    const double g2dq = timed([&] {
        for (int j = 0; j < 4; ++j) cudaMemcpy2DAsync(h[j], pitch, d, pitch, width, rows, cudaMemcpyDeviceToHost, st[j]);
    }) * 4;
    double gk[3];
    const int grids[3] = {148, 296, 1184};
    for (int g = 0; g < 3; ++g)
        gk[g] = timed([&] {
            void* hd = nullptr;
            cudaHostGetDevicePointer(&hd, h[k++ % nbuf], 0);
            store_rows<<<grids[g], 256, 0, st[0]>>>((const int4*)d, (int4*)hd, pitch / 16, width / 16, (int)rows);
        });
    // does the host row pitch matter? (rows start mid-page at the natural pitch of 480 000 B)
    const size_t pitches[4] = {480000, 480256, 483328, 524288};
    double gp[4];
    void* hbig = nullptr;
    CK(cudaHostAlloc(&hbig, rows * 524288, cudaHostAllocPortable));
    for (int i = 0; i < 4; ++i) {
        const size_t hp = pitches[i];
        gp[i] = timed([&] { cudaMemcpy2DAsync(hbig, hp, d, pitch, width, rows, cudaMemcpyDeviceToHost, st[0]); });
    }
    const double gp_dev = timed([&] { cudaMemcpy2DAsync(hbig, 483328, d, 483328, width, rows - 8, cudaMemcpyDeviceToHost, st[0]); });
    // packed destination: rows land back to back on the host (dst pitch = row width), the source keeps its pitch
    const double gpk = timed([&] { cudaMemcpy2DAsync(hbig, width, d, pitch, width, rows, cudaMemcpyDeviceToHost, st[0]); });
    int engines = 0;
    cudaDeviceGetAttribute(&engines, cudaDevAttrAsyncEngineCount, 0);
    std::printf("{\"what\": \"int16 hand-off of one receiver to pinned host memory, GB/s of the %zu demodulated bytes\", "
                "\"async_engines\": %d, \"memcpy_1d\": %.2f, \"memcpy_2d\": %.2f, \"memcpy_2d_2streams\": %.2f, "
                "\"memcpy_2d_4streams\": %.2f, \"memcpy_2d_4receivers_in_flight\": %.2f, "
                "\"kernel_zero_copy_148x256\": %.2f, \"kernel_zero_copy_296x256\": %.2f, \"kernel_zero_copy_1184x256\": %.2f, "
                "\"memcpy_2d_host_pitch_480000\": %.2f, \"memcpy_2d_host_pitch_480256\": %.2f, \"memcpy_2d_host_pitch_483328\": %.2f, "
                "\"memcpy_2d_host_pitch_524288\": %.2f, \"memcpy_2d_both_pitches_483328\": %.2f, \"memcpy_2d_packed_destination\": %.2f}\n",
                rows * width, engines, g1d, g2d, g2d2, g2d4, g2dq, gk[0], gk[1], gk[2], gp[0], gp[1], gp[2], gp[3], gp_dev, gpk);
    return 0;
}
