// Dynamic-range guard of the STFT mode (CWSL_MODE_STFT).
//
// The channelizer's error in a channel is the rounding noise of one float32 FFT that all channels of the receiver
// share: measured on B200 (tools/r2_probe.py, profiles/r2_guard_floor_run1.json) 0.6e-7 (median) ... 1.9e-7 (worst channel)
// of the rms of the WHOLE band on noise-like bands -- 10 butterfly stages + window + inter-pass twiddle, each ~0.5 ulp
// -- and up to 5.8e-7 (-125 dB) when one or two carriers hold nearly all of the band's power, whatever the channel
// itself holds. The direct form (source/SSBD.hpp:160-183, and the FAST kernel) keeps 15-17 dB more between a quiet
// channel and strong signals elsewhere in the band (2.5e-8 ... 1.0e-7). north_star asks for a residual >= 90 dB
// below the signal, so a channel whose own level is too close to the FFT's floor must not be taken from the FFT:
// kept channel segments lie >= -32 dB of the band, i.e. >= 93 dB above the worst floor measured.
//
// Per launch and per SEGMENT (fixed slot-relative spans of fast_seg_blocks() outputs, cwsl_kernels.hpp):
//   guard_band_power_kernel   P[s]  = mean |x|^2 of the segment's IQ                      (before the channelizer)
//   demod_chan_kernel         E[s][c] = sum y_c^2 (as an exact integer sum) and max|y_c|  (cwsl_chan.cu)
//   guard_select_kernel       channel c keeps the FFT result for segment s iff  E/n >= T^2 P; the others are
//                             listed, in channel order, as work items of <= 32 channels for
//   demod_fast_kernel<IND>    which recomputes them in the direct form and overwrites the float audio
//                             (cwsl_kernels.cu). max|x| of the kept channels is merged here, that of the recomputed
//                             ones by the FAST kernel, so prepareAudio (source/Instance.cpp:294-338) sees the maximum
//                             of exactly the samples that are handed on.
// Every statistic is order-independent (integer atomics, fixed-order reductions), so the decision -- hence every
// output byte -- is a function of the slot's IQ alone: not of chunking, CTA scheduling or channel order.
#include "cwsl_kernels.hpp"

#include <algorithm>

namespace cwsl {

namespace {

constexpr int kGuardThreads = 256;
constexpr int kPowerThreads = 1024;  // one CTA per segment: 24 064 samples at 192 kHz, ~24 per thread, 8 loads in flight
constexpr int kPowerUnroll = 8;

__global__ void __launch_bounds__(kPowerThreads) guard_band_power_kernel(DemodLaunch p, GuardLaunch g) {
    const uint32_t seg = blockIdx.x, t = threadIdx.x;
    if (seg == 0 && t == 0) *g.n_items = 0u;  // work list of this launch starts empty
    const uint32_t BS = p.block_size;
    const uint32_t k0 = p.b0 + seg * g.seg_blocks;
    const uint32_t k1 = min(k0 + g.seg_blocks, p.b1);
    const uint32_t n = (k1 - k0) * BS;  // complex samples of the segment
    const uint32_t row0 = (uint32_t)(((uint64_t)p.ring_off + k0) % p.ring_blocks);
    auto sample = [&](uint32_t i) {
        float2 x = make_float2(0.0f, 0.0f);
        if (i < n) x = __ldg(p.iq_ring + (size_t)((row0 + i / BS) % p.ring_blocks) * BS + (i % BS));
        return x;
    };
    float acc[kPowerUnroll];
#pragma unroll
    for (int u = 0; u < kPowerUnroll; ++u) acc[u] = 0.0f;
    for (uint32_t i = t; i < n; i += kPowerUnroll * kPowerThreads) {
        float2 x[kPowerUnroll];
#pragma unroll
        for (int u = 0; u < kPowerUnroll; ++u) x[u] = sample(i + u * kPowerThreads);
#pragma unroll
        for (int u = 0; u < kPowerUnroll; ++u) acc[u] = fmaf(x[u].y, x[u].y, fmaf(x[u].x, x[u].x, acc[u]));
    }
    // fixed-order reduction: the thread's partial sums, a shuffle tree inside the warp, the 32 warp sums in order
    double d = 0.0;
#pragma unroll
    for (int u = 0; u < kPowerUnroll; ++u) d += (double)acc[u];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) d += __shfl_down_sync(0xffffffffu, d, o);
    __shared__ double warp_sum[kPowerThreads / 32];
    if ((t & 31u) == 0) warp_sum[t >> 5] = d;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kPowerThreads / 32; ++w) s += warp_sum[w];
        const double P = n ? s / (double)n : 0.0;
        // octet contributions are sum_8 y^2 * scale, so that the keep-threshold E/n >= T^2 P reads sum >= 128 * n
        g.seg_scale[seg] = P > 0.0 ? (float)(128.0 / ((double)g.t2 * P)) : 0.0f;  // (inf for denormal-scale input: everything is kept)
    }
}

__global__ void __launch_bounds__(kGuardThreads) guard_select_kernel(DemodLaunch p, GuardLaunch g) {
    const uint32_t seg = blockIdx.x, t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t C = p.n_channels;
    const uint32_t hops = min(g.seg_blocks, p.b1 - p.b0 - seg * g.seg_blocks);
    const uint32_t thr = hops * 128u;
    const bool active = g.seg_scale[seg] > 0.0f;  // all-zero IQ: both forms give zeros, nothing to redo
    __shared__ uint32_t warp_cnt[kGuardThreads / 32];
    __shared__ uint32_t base;
    if (t == 0) base = 0u;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < C; c0 += kGuardThreads) {
        const uint32_t c = c0 + t;
        const bool valid = c < C;
        unsigned e = 0xffffffffu, m = 0u;
        if (valid) {
            const size_t i = (size_t)seg * g.stat_stride + c;
            e = g.seg_energy[i];
            m = g.seg_max[i];
            g.seg_energy[i] = 0u;  // ready for the next launch
            g.seg_max[i] = 0u;
        }
        const bool sel = valid && active && e < thr;
        if (valid && !sel && m != 0u) atomicMax(p.maxbits + c, m);
        const unsigned ballot = __ballot_sync(0xffffffffu, sel);
        if (lane == 0) warp_cnt[warp] = __popc(ballot);
        __syncthreads();
        uint32_t off = base;
        for (uint32_t w = 0; w < warp; ++w) off += warp_cnt[w];
        if (sel) g.sel[(size_t)seg * g.sel_stride + off + __popc(ballot & ((1u << lane) - 1u))] = c;
        __syncthreads();
        if (t == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < kGuardThreads / 32; ++w) tot += warp_cnt[w];
            base += tot;
        }
        __syncthreads();
    }
    if (t == 0) {
        const uint32_t nsel = base;
        if (g.counters) {
            atomicAdd(g.counters + 0, (unsigned long long)C);
            if (nsel) atomicAdd(g.counters + 1, (unsigned long long)nsel);
        }
        if (nsel) {
            const uint32_t n = (nsel + 31u) / 32u;
            const uint32_t first = atomicAdd(g.n_items, n);
            for (uint32_t j = 0; j < n; ++j) {
                GuardItem it;
                it.seg = seg;
                it.first = 32u * j;
                it.count = min(32u, nsel - 32u * j);
                it.pad = 0u;
                g.items[first + j] = it;
            }
        }
    }
}

}  // namespace

cudaError_t launch_guard_band_power(const DemodLaunch& p, const GuardLaunch& g, cudaStream_t s) {
    if (p.b1 <= p.b0 || g.seg_blocks == 0) return cudaSuccess;
    if (p.b0 % g.seg_blocks != 0) return cudaErrorInvalidValue;
    const uint32_t n_seg = (p.b1 - p.b0 + g.seg_blocks - 1) / g.seg_blocks;
    guard_band_power_kernel<<<n_seg, kPowerThreads, 0, s>>>(p, g);
    return cudaGetLastError();
}

cudaError_t launch_guard_select(const DemodLaunch& p, const GuardLaunch& g, cudaStream_t s) {
    if (p.b1 <= p.b0 || g.seg_blocks == 0 || p.n_channels == 0) return cudaSuccess;
    const uint32_t n_seg = (p.b1 - p.b0 + g.seg_blocks - 1) / g.seg_blocks;
    guard_select_kernel<<<n_seg, kGuardThreads, 0, s>>>(p, g);
    return cudaGetLastError();
}

}  // namespace cwsl
