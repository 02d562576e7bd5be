// sm_100a kernels of the CWSL_DIGI receive front-end: NCO mix -> LowPass FIR -> decimate ->
// Weaver SSB demod (source/SSBD.hpp:128-183) and normalise -> int16 (source/Instance.cpp:238-241,
// :294-338), batched over every decoder channel of a receiver.
//
// Math. For channel c, SSBD block k = BS input samples:
//   S[k][n]  = sum_{m<BS} (x[BS*k+m] * tone_c[m]) * h[BS*n+m]            n = 0..31
//   y_c[b]   = sum_{n=0..31, k=b-31+n>=0} S[k][n] * P_c[k]               (ascending n)
//   audio[b] = {+Re, -Im*sign, -Re, +Im*sign}[b & 3] of y_c[b]
// P_c[k] = phase_inc_c^k by the reference's float recurrence (phase table, built once per channel).
//
// Kernels:
//   demod_fast_kernel        FAST mode. One thread per R consecutive blocks; the IQ rows of a tile are
//                            staged into shared memory by the TMA bulk-copy engine (cp.async.bulk +
//                            mbarrier) and re-used by all channels the CTA walks; mix and FIR are packed
//                            fma.rn.f32x2 (FFMA2) with the symmetric taps folded and every coefficient an
//                            FFMA2 immediate, so the inner product issues no load at all; partial sums are
//                            exchanged through shared memory and carried across tiles; Weaver select,
//                            max|x| and the float audio store are fused in the epilogue.
//   demod_exact_tiled_kernel EXACT mode. The reference's loop structure with one thread per block, every
//                            float operation unfused in the reference's order -> bit-identical.
//   demod_exact_kernel       EXACT mode, independent one-thread-per-output implementation (cross-check).
//   quantise_kernel          prepareAudio + int16 conversion for all channels.
//   phase_table_kernel       the float phase recurrence, one thread per table.
#include "cwsl_kernels.hpp"
#include "cwsl_ptx.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <set>
#include <string>
#include <utility>

namespace cwsl {

// ------------------------------------------------------------------------------------------
// Constant-bank taps h[BS*n+m] for the gather EXACT kernel, one array per supported block size
// (Fs = 12000*BS), so receivers of different rates coexist. (The tiled EXACT kernel and the FAST kernel
// use the same values baked into the instruction stream, see cwsl_taps_baked.inc.)
// ------------------------------------------------------------------------------------------
__constant__ float c_h16[512];
__constant__ float c_h8[256];
__constant__ float c_h4[128];

template <int BS>
__device__ __forceinline__ float tap_nat(int i) {
    if constexpr (BS == 16) return c_h16[i];
    else if constexpr (BS == 8) return c_h8[i];
    else return c_h4[i];
}
cudaError_t upload_taps(uint32_t block_size, const float* taps) {
    const uint32_t n = 32 * block_size;
    switch (block_size) {
        case 16: return cudaMemcpyToSymbol(c_h16, taps, n * sizeof(float));
        case 8: return cudaMemcpyToSymbol(c_h8, taps, n * sizeof(float));
        case 4: return cudaMemcpyToSymbol(c_h4, taps, n * sizeof(float));
        default: return cudaErrorInvalidValue;
    }
}

// Compile-time copy of the taps with the FIR row update unrolled around them (FFMA2 immediates).
template <int BS>
struct FirBlock;   // all 32 tap rows of one SSBD block (symmetric/folded form for BS >= 8), FAST mode
template <int BS>
struct ExactRows;  // 8 tap rows at a time, unfused mul/add in the reference's order, EXACT mode
#include "cwsl_taps_baked.inc"

const float* baked_taps_transposed(uint32_t block_size) {
    switch (block_size) {
        case 16: return kTapsT16;
        case 8: return kTapsT8;
        case 4: return kTapsT4;
        default: return nullptr;
    }
}

// Weaver select of SSBD::Iterate (source/SSBD.hpp:132-135); sign flips are exact.
__device__ __forceinline__ float weaver(float re, float im, uint32_t b, float sign) {
    switch (b & 3u) {
        case 0: return re;
        case 1: return __fmul_rn(-im, sign);
        case 2: return -re;
        default: return __fmul_rn(im, sign);
    }
}

// ------------------------------------------------------------------------------------------
// Phase tables: P[k] = phase_inc^k, the float recurrence of source/SSBD.hpp:174 with
// libstdc++'s complex multiply (ac-bd, ad+bc), each operation rounded separately (no FMA).
// Sequential per table by construction (a parallel scan would change the rounding).
// ------------------------------------------------------------------------------------------
__global__ void phase_table_kernel(const float2* __restrict__ inc, float2* const* __restrict__ tables, uint32_t n,
                                   uint32_t length) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float ir = inc[i].x, ii = inc[i].y;
    float2* __restrict__ out = tables[i];
    float pr = 1.0f, pi = 0.0f;
    for (uint32_t k = 0; k < length; ++k) {
        out[k] = make_float2(pr, pi);
        const float a = __fmul_rn(pr, ir), b = __fmul_rn(pi, ii);
        const float c = __fmul_rn(pr, ii), d = __fmul_rn(pi, ir);
        pr = __fsub_rn(a, b);
        pi = __fadd_rn(c, d);
    }
}

cudaError_t launch_phase_tables(const float2* phase_inc, float2* const* tables, uint32_t n, uint32_t length,
                                cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    phase_table_kernel<<<(n + 31) / 32, 32, 0, s>>>(phase_inc, tables, n, length);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// EXACT demodulator: gather form, one thread per (channel, output sample); unfused float ops in
// the reference's order (source/SSBD.hpp:164-181) -> bit-identical to the reference chain.
// ------------------------------------------------------------------------------------------
template <int BS>
__global__ void __launch_bounds__(128) demod_exact_kernel(DemodLaunch p) {
    const uint32_t c = blockIdx.y;
    const uint32_t b = p.b0 + blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float2 s_tone[BS];
    if (threadIdx.x < BS) s_tone[threadIdx.x] = p.tone[(size_t)c * BS + threadIdx.x];
    __syncthreads();
    float out = 0.0f;
    const bool live = b < p.b1;
    if (live) {
        const float2* __restrict__ P = p.phase[c];
        float wr = 0.0f, wi = 0.0f;
#pragma unroll 1
        for (int n = 0; n < 32; ++n) {
            const int64_t k = (int64_t)b - 31 + n;
            if (k < 0) continue;  // fresh SSBD: zero history (source/Instance.cpp:251)
            const uint32_t row = (uint32_t)((p.ring_off + (uint64_t)k) % p.ring_blocks);
            const float4* __restrict__ x = reinterpret_cast<const float4*>(p.iq_ring + (size_t)row * BS);
            float sr = 0.0f, si = 0.0f;
#pragma unroll
            for (int m2 = 0; m2 < BS / 2; ++m2) {
                const float4 xx = __ldg(x + m2);
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int m = 2 * m2 + s;
                    const float xr = s ? xx.z : xx.x, xi = s ? xx.w : xx.y;
                    const float tr = s_tone[m].x, ti = s_tone[m].y;
                    const float vr = __fsub_rn(__fmul_rn(xr, tr), __fmul_rn(xi, ti));  // in[m]*tone[m]
                    const float vi = __fadd_rn(__fmul_rn(xr, ti), __fmul_rn(xi, tr));
                    const float h = tap_nat<BS>(n * BS + m);
                    sr = __fadd_rn(sr, __fmul_rn(vr, h));  // sum += (..)*filter[m+n*BlockSize]
                    si = __fadd_rn(si, __fmul_rn(vi, h));
                }
            }
            const float2 ph = P[k];
            wr = __fadd_rn(wr, __fsub_rn(__fmul_rn(sr, ph.x), __fmul_rn(si, ph.y)));  // workspace += sum*phase
            wi = __fadd_rn(wi, __fadd_rn(__fmul_rn(sr, ph.y), __fmul_rn(si, ph.x)));
        }
        out = weaver(wr, wi, b, p.sign[c]);
        p.audio[(size_t)c * p.af_stride + b] = out;
    }
    atomic_max_abs(p.maxbits + c, live ? fabsf(out) : 0.0f);
}

// ------------------------------------------------------------------------------------------
// FAST demodulator.
//
// CTA = NT threads; thread t owns R consecutive SSBD blocks kbase = kt0 + t*R .. +R-1 ("row").
// A CTA owns a time SEGMENT of `tiles_per_seg` tiles (tile = NT*R blocks) and a group of <= 32
// channels. It walks the tiles forward in time; per tile the rows are brought into shared memory
// by per-row TMA bulk copies (row stride padded by 16 B so the per-thread LDS.128 of "my row" are
// bank-conflict free) and stay there while the CTA walks its channels. Per channel every thread:
//   mixes its R blocks with w[m] = tone_c[m]*P_c[k] (packed complex multiply),
//   accumulates acc[31-n] += v[m]*h[16n+m] for all 32 tap rows (FirBlock: folded symmetric taps,
//   FFMA2 with immediate taps) -> 31+R partial sums indexed by output offset,
//   hands acc[R..30+R] to the later threads through shared memory and finishes its own R outputs
//   with what the (<= 8) earlier threads handed over.
// What the last 32/R threads of a tile owe to the NEXT tile's first 31 outputs is reduced to 31
// partial sums per channel ("carry") kept in shared memory, so only the first tile of a segment
// recomputes a 32-block overlap.
// ------------------------------------------------------------------------------------------
#ifndef CWSL_FAST_RUNROLL
#define CWSL_FAST_RUNROLL 4  // block loop fully unrolled: +3.7 % over the rolled loop (which ptxas unrolls by two),
#endif                       // no register-shift MOVs, static phase selects; 41 KB of code, no I-cache penalty measured
constexpr int kFastRUnroll = CWSL_FAST_RUNROLL;
constexpr int kFastGMax = 32;  // channels walked per CTA (tone tables, phase pointers, carries staged in smem)

template <int BS, int R, int NT, int G = kFastGMax>
struct FastCfg {
    static constexpr int kG = G;  // channels walked per CTA
    static constexpr int kRowBytes = R * BS * 8;
    static constexpr int kRowStride = kRowBytes + 16;
    static constexpr int kHaloT = 32 / R;
    static constexpr int kTile = NT * R;  // blocks per tile
    static constexpr int kNE = 31;
    static constexpr size_t kXBytes = (size_t)NT * kRowStride;
    static constexpr size_t kEBytes = (size_t)NT * kNE * 8;
    static constexpr size_t kOBytes = (size_t)NT * R * 8;
    static constexpr size_t kToneBytes = (size_t)G * BS * 8;
    static constexpr size_t kCarryBytes = (size_t)G * 31 * 8;
    // 2 CTAs/SM need 2*(kSmem + 1 KB reserved) <= 228 KB: 115 472 B for <16,4,128>
    // channel index of every walked channel (direct: c0 + i; guard: from the selection list), sideband in bit 31.
    // (Not one byte more than this: 2 CTAs/SM is the whole point of the shape, and 128 B more lose the second CTA --
    // measured 6.06 instead of 4.85 ms per 1024-channel FT8 slot.)
    static constexpr size_t kIdxBytes = (size_t)G * 4;
    static constexpr size_t kSmem = kXBytes + kEBytes + kOBytes + kToneBytes + kCarryBytes + kIdxBytes + 16;
};

// IND = false: grid (segments, channel groups), one work item per CTA. IND = true (dynamic-range guard of the STFT
// mode): persistent 1-D grid, CTA i walks items i, i + gridDim.x, ... of a device-resident list; an item is <= 32
// channels of one segment, named by index in the segment's selection list.
template <int BS, int R, int NT, int CTAS, bool PF, bool IND, int G = kFastGMax>
__global__ void __launch_bounds__(NT, CTAS)
    demod_fast_kernel(DemodLaunch p, uint32_t ch_per_cta, uint32_t tiles_per_seg, FastIndirect ind) {
    using Cfg = FastCfg<BS, R, NT, G>;
    static_assert(32 % R == 0 && (R == 2 || R == 4), "R must be 2 or 4");
    static_assert(NT >= 32 && Cfg::kHaloT <= NT, "tile too small");
    static_assert(Cfg::kTile == (int)kFastTile, "segment geometry is part of the launch interface");
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* xs = smem;
    float2* E = reinterpret_cast<float2*>(smem + Cfg::kXBytes);
    float2* O = reinterpret_cast<float2*>(smem + Cfg::kXBytes + Cfg::kEBytes);
    float4* tone_s = reinterpret_cast<float4*>(smem + Cfg::kXBytes + Cfg::kEBytes + Cfg::kOBytes);
    float2* carry_s = reinterpret_cast<float2*>(smem + Cfg::kXBytes + Cfg::kEBytes + Cfg::kOBytes + Cfg::kToneBytes);
    uint32_t* cidx_s = reinterpret_cast<uint32_t*>(smem + Cfg::kXBytes + Cfg::kEBytes + Cfg::kOBytes + Cfg::kToneBytes +
                                                   Cfg::kCarryBytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Cfg::kXBytes + Cfg::kEBytes + Cfg::kOBytes + Cfg::kToneBytes +
                                                Cfg::kCarryBytes + Cfg::kIdxBytes);
    constexpr uint32_t kIdxMask = 0x7fffffffu;

    const int t = threadIdx.x;
    const uint32_t bar_a = smem_u32(bar);
    if (t == 0) {
        mbar_init(bar_a, 1);
        fence_mbar_init();
    }
    uint32_t n_wait = 0;  // mbarrier phases consumed so far (one per staged tile)
    const uint32_t n_items = IND ? __ldg(ind.n_items) : 1u;
    for (uint32_t item = IND ? blockIdx.x : 0u; item < n_items; item += IND ? gridDim.x : 1u) {
    uint32_t seg, nch;
    if constexpr (IND) {
        const GuardItem it = ind.items[item];
        seg = it.seg;
        nch = it.count;
        __syncthreads();  // (first item: the mbarrier init; later items: nothing reads the staged tables any more)
        if ((uint32_t)t < nch) {
            const uint32_t c = __ldg(ind.sel + (size_t)seg * ind.sel_stride + it.first + t);
            cidx_s[t] = c | (p.sign[c] < 0.0f ? 0x80000000u : 0u);
        }
    } else {
        seg = blockIdx.x;
        const uint32_t c0 = blockIdx.y * ch_per_cta;
        nch = min(ch_per_cta, p.n_channels - c0);
        if ((uint32_t)t < nch) cidx_s[t] = (c0 + t) | (p.sign[c0 + t] < 0.0f ? 0x80000000u : 0u);
    }
    // segment: outputs [seg_b0, seg_b1); its first tile starts 32 blocks early (overlap, no output there)
    const int64_t seg_out = (int64_t)tiles_per_seg * Cfg::kTile - 32;
    const int64_t seg_b0 = (int64_t)p.b0 + (int64_t)seg * seg_out;
    const int64_t seg_b1 = min(seg_b0 + seg_out, (int64_t)p.b1);
    __syncthreads();
    for (uint32_t i = t; i < nch * (BS / 2); i += NT)
        tone_s[i] = reinterpret_cast<const float4*>(p.tone)[(size_t)(cidx_s[i / (BS / 2)] & kIdxMask) * (BS / 2) + i % (BS / 2)];
    __syncthreads();

    const float4* __restrict__ xrow = reinterpret_cast<const float4*>(xs + (size_t)t * Cfg::kRowStride);
    float2* __restrict__ Et = E + (size_t)t * Cfg::kNE;

    for (uint32_t tile = 0; tile < tiles_per_seg; ++tile) {
        const int64_t kt0 = seg_b0 - 32 + (int64_t)tile * Cfg::kTile;
        if (tile > 0 && kt0 >= seg_b1) break;  // nothing left to output (tiles after the first start AT kt0)
        const int64_t kbase = kt0 + (int64_t)t * R;
        const bool row_valid = kbase >= 0 && kbase + R <= (int64_t)p.b1;
        // threads whose outputs are written: inside the segment; the overlap rows of the first tile are not
        const bool writes = row_valid && kbase >= seg_b0 && kbase < seg_b1;
        const bool first_tile = tile == 0;

        // ---- stage this tile's IQ rows (TMA bulk copies, one per thread row) ----
        if (t == 0) {
            int64_t lo = kt0 >= 0 ? 0 : (-kt0 + R - 1) / R;
            int64_t hi = ((int64_t)p.b1 - kt0) / R;
            if (hi > NT) hi = NT;
            const uint32_t rows = hi > lo ? (uint32_t)(hi - lo) : 0u;
            mbar_arrive_expect_tx(bar_a, rows * Cfg::kRowBytes);
        }
        if (row_valid) {
            const uint32_t row = (uint32_t)((p.ring_off + (uint64_t)kbase) % p.ring_blocks);
            tma_bulk_g2s(smem_u32(xs + (size_t)t * Cfg::kRowStride), p.iq_ring + (size_t)row * BS, Cfg::kRowBytes,
                         bar_a);
        }
        // Phase-table residency (STFT guard): replay this tile's phases from the anchors into the CTA's scratch slice.
        // (channel, anchor interval) tasks of <= 128 sequential steps each, <= 5 intervals per tile; the previous
        // tile's readers are past the __syncthreads that ends every channel.
        const float2* scr = nullptr;
        if constexpr (IND) {
            if (ind.anchors) {
                float2* scr_w = ind.scratch + (size_t)blockIdx.x * G * Cfg::kTile;
                const int64_t k_lo = kt0 > 0 ? kt0 : 0, k_hi = min(kt0 + (int64_t)Cfg::kTile, (int64_t)p.b1);
                if (k_hi > k_lo) {
                    const uint32_t a_first = (uint32_t)(k_lo / kChanAnchorHops);
                    const uint32_t n_int = (uint32_t)((k_hi - 1) / kChanAnchorHops) - a_first + 1u;
                    for (uint32_t task = t; task < nch * n_int; task += NT) {
                        const uint32_t ci = task / n_int, ai = a_first + task % n_int;
                        const uint32_t c = cidx_s[ci] & kIdxMask;
                        float2 P = __ldg(ind.anchors + (size_t)ai * ind.anchor_stride + c);
                        const float2 inc = __ldg(ind.pinc + c);
                        int64_t k = (int64_t)ai * kChanAnchorHops;
                        const int64_t k_end = min(k + (int64_t)kChanAnchorHops, k_hi);
                        float2* dst = scr_w + (size_t)ci * Cfg::kTile - kt0;
                        for (; k < k_end; ++k) {
                            if (k >= k_lo) dst[k] = P;
                            const float x = __fmul_rn(P.x, inc.x), y = __fmul_rn(P.y, inc.y);  // phase_table_kernel's ops
                            const float z = __fmul_rn(P.x, inc.y), w = __fmul_rn(P.y, inc.x);
                            P = make_float2(__fsub_rn(x, y), __fadd_rn(z, w));
                        }
                    }
                }
                __syncthreads();
                scr = scr_w - kt0;  // scr[ci * kTile + k]
            }
        }
        // phase values of my R blocks for the first channel (later channels are prefetched one ahead:
        // the tables live in HBM, ~1 us away at 2 warps per scheduler)
        float4 Pnext[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) Pnext[i] = make_float4(1.f, 0.f, 1.f, 0.f);
        if (row_valid) {
            if (IND && scr) {
                const float4* pp = reinterpret_cast<const float4*>(scr + kbase);
#pragma unroll
                for (int i = 0; i < R / 2; ++i) Pnext[i] = __ldcg(pp + i);  // (written by this CTA: not the read-only path)
            } else {
                const float4* pp = reinterpret_cast<const float4*>(p.phase[cidx_s[0] & kIdxMask] + kbase);
#pragma unroll
                for (int i = 0; i < R / 2; ++i) Pnext[i] = __ldg(pp + i);
            }
        }
        mbar_wait(bar_a, n_wait & 1u);
        ++n_wait;

        // IQ samples and tone entries of the NEXT block to be mixed live in registers: they are dead as soon
        // as the block is mixed, so the next block's are fetched right then and their LDS latency hides under
        // the ~340 FFMA2 of the FIR (the rows are the same for every channel, only the tones change)
        float4 xq[BS / 2], tq[BS / 2];
        if (PF && row_valid) {
#pragma unroll
            for (int m2 = 0; m2 < BS / 2; ++m2) {
                xq[m2] = xrow[m2];
                tq[m2] = tone_s[m2];
            }
        }

        for (uint32_t ci = 0; ci < nch; ++ci) {
            const uint32_t c_sb = cidx_s[ci], c = c_sb & kIdxMask;
            // acc[i] = partial sum for output offset (blocks done so far) + i
            float2 acc[32];
            float2 own[R];  // finished-as-far-as-I-am-concerned sums of my own R outputs
#pragma unroll
            for (int j = 0; j < R; ++j) own[j] = make_float2(0.0f, 0.0f);
            if (!row_valid || kFastRUnroll < R) {  // (a fully unrolled first block assigns every accumulator)
#pragma unroll
                for (int o = 0; o < 32; ++o) acc[o] = make_float2(0.0f, 0.0f);
            }

            if (row_valid) {
                float4 Pcur[R / 2];
#pragma unroll
                for (int i = 0; i < R / 2; ++i) Pcur[i] = Pnext[i];
                if (ci + 1 < nch) {
                    if (IND && scr) {
                        const float4* pp = reinterpret_cast<const float4*>(scr + (size_t)(ci + 1) * Cfg::kTile + kbase);
#pragma unroll
                        for (int i = 0; i < R / 2; ++i) Pnext[i] = __ldcg(pp + i);
                    } else {
                        const float4* pp = reinterpret_cast<const float4*>(p.phase[cidx_s[ci + 1] & kIdxMask] + kbase);
#pragma unroll
                        for (int i = 0; i < R / 2; ++i) Pnext[i] = __ldg(pp + i);
                    }
                }
#pragma unroll(kFastRUnroll)
                for (int r = 0; r < R; ++r) {
                    float4 Pq = Pcur[0];
#pragma unroll
                    for (int i = 1; i < R / 2; ++i) Pq = (r >> 1) == i ? Pcur[i] : Pq;
                    const float2 Pk = (r & 1) ? make_float2(Pq.z, Pq.w) : make_float2(Pq.x, Pq.y);
                    // mix: v[m] = x[m] * (tone[m] * P[k]); (a+ib)(c+id) = a*(c,d) + b*(-d,c)
                    float2 v[BS];
#pragma unroll
                    for (int m2 = 0; m2 < BS / 2; ++m2) {
                        const float4 xx = PF ? xq[m2] : xrow[r * (BS / 2) + m2];                    // two IQ samples
                        const float4 tt = PF ? tq[m2] : tone_s[(size_t)ci * (BS / 2) + m2];        // their two tone entries
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            const float2 tn_m = s ? make_float2(tt.z, tt.w) : make_float2(tt.x, tt.y);
                            float2 w = fmul2(tn_m, bc(Pk.x));
                            w = ffma2(make_float2(-tn_m.y, tn_m.x), bc(Pk.y), w);
                            const float xr = s ? xx.z : xx.x, xi = s ? xx.w : xx.y;
                            float2 vv = fmul2(w, bc(xr));
                            v[2 * m2 + s] = ffma2(make_float2(-w.y, w.x), bc(xi), vv);
                        }
                    }
                    // fetch the next block's samples/tones (next r, or block 0 of the next channel)
                    if constexpr (PF) {
                        const int rn = (r + 1 == R) ? 0 : r + 1;
                        const uint32_t cn = (r + 1 == R) ? min(ci + 1, nch - 1) : ci;
                        const float4* __restrict__ xn = xrow + rn * (BS / 2);
                        const float4* __restrict__ tn = tone_s + (size_t)cn * (BS / 2);
#pragma unroll
                        for (int m2 = 0; m2 < BS / 2; ++m2) {
                            xq[m2] = xn[m2];
                            tq[m2] = tn[m2];
                        }
                    }
                    // FIR: tap row n of this block feeds output offset 31-n
                    // The first block of a row starts the accumulators: assign instead of add onto zeros -- only when
                    // the loop is fully unrolled (r static); in a rolled loop the branch stops ptxas from unrolling.
                    if (kFastRUnroll >= R && r == 0)
                        FirBlock<BS>::template apply<true>(v, acc);
                    else
                        FirBlock<BS>::template apply<false>(v, acc);
                    // offset 0 is complete as far as this thread is concerned; slide the window
                    if constexpr (kFastRUnroll >= R) own[r] = acc[0];  // r is static: stays in a register
                    else O[r * NT + t] = acc[0];
#pragma unroll
                    for (int o = 0; o < 31; ++o) acc[o] = acc[o + 1];  // acc[31] is assigned by the next block
                    acc[31] = make_float2(0.0f, 0.0f);
                }
            } else if (kFastRUnroll < R) {
#pragma unroll
                for (int r = 0; r < R; ++r) O[r * NT + t] = make_float2(0.0f, 0.0f);
            }

            // ---- exchange partial sums: acc[0..30] are offsets R..R+30 ----
#pragma unroll
            for (int o = 0; o < 31; ++o) Et[o] = acc[o];
            __syncthreads();
            if constexpr (kFastRUnroll < R) {
#pragma unroll
                for (int j = 0; j < R; ++j) own[j] = O[j * NT + t];
            }
#pragma unroll
            for (int i = 1; i * R <= 30 + R; ++i) {
                if (t - i >= 0) {
                    const float2* __restrict__ Ei = E + (size_t)(t - i) * Cfg::kNE;
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const int o = j + i * R;
                        if (o <= 30 + R) own[j] = fadd2(own[j], Ei[o - R]);
                    }
                }
            }
            // what the previous tile owes to my outputs (first 31 output offsets of the tile)
            if (!first_tile && t < Cfg::kHaloT) {
#pragma unroll
                for (int j = 0; j < R; ++j)
                    if (t * R + j < 31) own[j] = fadd2(own[j], carry_s[ci * 31 + t * R + j]);
            }
            __syncwarp();
            // what this tile owes to the next one: for output offset j' past the tile end, the hand-overs
            // of the last threads that the (virtual) thread NT + j'/R would have collected
            if (t < 31) {
                const int q = t / R, j = t % R;
                float2 cs = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int i = 1; i * R <= 30 + R; ++i) {
                    const int o = j + i * R;
                    if (i > q && o <= 30 + R) cs = fadd2(cs, E[(size_t)(NT + q - i) * Cfg::kNE + o - R]);
                }
                carry_s[ci * 31 + t] = cs;
            }

            // ---- epilogue: Weaver select, float audio store, max|x| ----
            float lmax = 0.0f;
            if (writes) {
                const float sign = (c_sb >> 31) ? -1.0f : 1.0f;
                float o4[R];
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    o4[j] = weaver(own[j].x, own[j].y, (uint32_t)(kbase + j), sign);
                    lmax = fmaxf(lmax, fabsf(o4[j]));
                }
                float* dst = p.audio + (size_t)c * p.af_stride + kbase;
                if constexpr (R == 4)
                    *reinterpret_cast<float4*>(dst) = make_float4(o4[0], o4[1], o4[2], o4[3]);
                else
                    *reinterpret_cast<float2*>(dst) = make_float2(o4[0], o4[1]);
            }
            atomic_max_abs(p.maxbits + c, lmax);
            __syncthreads();  // E/O are rewritten by the next channel, the IQ rows by the next tile
        }
    }
    }  // items
}

// Opt the kernel in to its dynamic shared memory size (per device: function attributes are per context) and
// return the SM count of the current device. Cheap enough to call on every launch; thread-safe.
static cudaError_t prepare_kernel(const void* kern, int smem_bytes, int* sms) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    static int sm_count[64] = {0};
    std::lock_guard<std::mutex> lk(mu);
    if (!done.count({kern, dev})) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return e;
        done.insert({kern, dev});
    }
    if (dev >= 0 && dev < 64) {
        if (sm_count[dev] == 0) {
            int n = 0;
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
            sm_count[dev] = n > 0 ? n : 148;
        }
        *sms = sm_count[dev];
    } else {
        *sms = 148;
    }
    return cudaSuccess;
}

// Tiles per segment: trade the 32-block overlap paid once per segment against the tail of the last
// wave of CTAs. Cost model in block units, minimised over L.
static uint32_t choose_tiles_per_seg(uint32_t n_out, uint32_t tile, uint32_t ch_groups, uint32_t slots,
                                     uint32_t overlap = 32) {
    // (EXACT mode only: its result is bit-identical however it is segmented.) CWSL_TILES_PER_SEG=<n> pins the
    // segment length (tests and compute-sanitizer runs use it to force the cross-tile carry path on small inputs)
    static const uint32_t forced = [] {
        const char* e = std::getenv("CWSL_TILES_PER_SEG");
        return e ? (uint32_t)std::strtoul(e, nullptr, 10) : 0u;
    }();
    if (forced > 0) return forced;
    uint32_t best_l = 1;
    double best = 1e300;
    const uint32_t max_l = std::max<uint32_t>(1, std::min<uint32_t>(64, (n_out + tile - 1) / tile + 1));
    for (uint32_t l = 1; l <= max_l; ++l) {
        const uint64_t seg_out = (uint64_t)l * tile - overlap;
        const uint64_t nseg = (n_out + seg_out - 1) / seg_out;
        const uint64_t ctas = nseg * ch_groups;
        const uint64_t waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * l * tile;  // every wave lasts as long as a full segment
        if (cost < best * 0.999) {
            best = cost;
            best_l = l;
        }
    }
    return best_l;
}

// ------------------------------------------------------------------------------------------
// EXACT demodulator, production form ("scatter"): the reference's own loop structure
// (source/SSBD.hpp:160-183) with one thread per SSBD block.
//   thread k: v[m] = in[m]*tone[m]                    (unfused complex multiply)
//             S[n] = sum_m v[m]*h[16n+m], m ascending (mul.rn.f32x2 + add.rn.f32x2, never contracted)
//             D[k][n] = S[n]*phase_k                  (unfused complex multiply)
//   thread b: y[b] = sum_{n ascending} D[b-31+n][n]   starting from +0
// which is exactly the order in which the reference's circular workspace receives its terms, so the
// result is bit-identical. Packed f32x2 instructions are two independent IEEE operations (re, im).
// The CTA walks its tiles forward in time; the ascending-n prefix sums that the last 31 blocks of
// a tile owe to the next tile's first 31 outputs are carried in shared memory (a prefix of the same
// ordered sum, hence still bit-identical).
// ------------------------------------------------------------------------------------------
template <int BS, int NT, int G>  // G = channels walked per CTA
struct ExactCfg {
    static constexpr int kRowBytes = BS * 8;
    static constexpr int kRowStride = kRowBytes + 16;   // conflict-free per-thread LDS.128
    static constexpr int kDStride = 33;                 // float2 per D row (32 + 1 pad: conflict-free both ways)
    static constexpr size_t kXBytes = (size_t)NT * kRowStride;
    static constexpr int kZeroRows = 31;                // D rows -31..-1: fresh-SSBD / previous-tile history, all +0
    static constexpr size_t kDBytes = (size_t)(NT + kZeroRows) * kDStride * 8;
    static constexpr size_t kToneBytes = (size_t)G * BS * 8;
    static constexpr size_t kCarryBytes = (size_t)G * 31 * 8;
    static constexpr size_t kSmem = kXBytes + kDBytes + kToneBytes + kCarryBytes + 16;
};

// (a + ib)(c + id) = (ac - bd, ad + bc), every product and the add/sub rounded separately: libstdc++'s
// complex multiply as compiled with -ffp-contract=off
__device__ __forceinline__ float2 cmul_unfused(float2 x, float2 y) {
    const float2 p = fmul2(y, bc(x.x));                      // (a*c, a*d)
    const float2 q = fmul2(make_float2(y.y, y.x), bc(x.y));  // (b*d, b*c)
    return make_float2(__fsub_rn(p.x, q.x), __fadd_rn(p.y, q.y));  // (ac - bd, ad + bc), scalar: see add_unfused
}

template <int BS, int NT, int CTAS, int G>
__global__ void __launch_bounds__(NT, CTAS)
    demod_exact_tiled_kernel(DemodLaunch p, uint32_t ch_per_cta, uint32_t tiles_per_seg) {
    using Cfg = ExactCfg<BS, NT, G>;
    static_assert(NT >= 62, "the carry loop assumes its source rows are never before the slot");
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* xs = smem;
    float2* D0 = reinterpret_cast<float2*>(smem + Cfg::kXBytes);
    float2* D = D0 + (size_t)Cfg::kZeroRows * Cfg::kDStride;  // row 0 of the tile; rows -31..-1 stay +0 forever
    float2* tone_s = reinterpret_cast<float2*>(smem + Cfg::kXBytes + Cfg::kDBytes);
    float2* carry_s = reinterpret_cast<float2*>(smem + Cfg::kXBytes + Cfg::kDBytes + Cfg::kToneBytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Cfg::kXBytes + Cfg::kDBytes + Cfg::kToneBytes + Cfg::kCarryBytes);

    const int t = threadIdx.x;
    const uint32_t c0 = blockIdx.y * ch_per_cta;
    const uint32_t nch = min(ch_per_cta, p.n_channels - c0);
    const float2* const* __restrict__ phase_g = p.phase + c0;
    // segment of outputs [seg_b0, seg_b1); its first tile starts 31 blocks early (history, no output)
    const int64_t seg_out = (int64_t)tiles_per_seg * NT - 31;
    const int64_t seg_b0 = (int64_t)p.b0 + (int64_t)blockIdx.x * seg_out;
    const int64_t seg_b1 = min(seg_b0 + seg_out, (int64_t)p.b1);

    const uint32_t bar_a = smem_u32(bar);
    if (t == 0) {
        mbar_init(bar_a, 1);
        fence_mbar_init();
    }
    for (uint32_t i = t; i < nch * BS; i += NT) tone_s[i] = p.tone[(size_t)c0 * BS + i];
    for (int i = t; i < Cfg::kZeroRows * Cfg::kDStride; i += NT) D0[i] = make_float2(0.0f, 0.0f);
    __syncthreads();

    const float4* __restrict__ xrow = reinterpret_cast<const float4*>(xs + (size_t)t * Cfg::kRowStride);
    float2* __restrict__ Dt = D + (size_t)t * Cfg::kDStride;

    for (uint32_t tile = 0; tile < tiles_per_seg; ++tile) {
        const int64_t kt0 = seg_b0 - 31 + (int64_t)tile * NT;
        if (tile > 0 && kt0 >= seg_b1) break;  // tiles after the first output from kt0 on
        const int64_t k = kt0 + t;  // my block = my output
        const bool row_valid = k >= 0 && k < (int64_t)p.b1;
        const bool writes = row_valid && k >= seg_b0 && k < seg_b1;
        const bool first_tile = tile == 0;

        if (t == 0) {
            const int64_t lo = kt0 >= 0 ? 0 : -kt0;
            int64_t hi = (int64_t)p.b1 - kt0;
            if (hi > NT) hi = NT;
            const uint32_t rows = hi > lo ? (uint32_t)(hi - lo) : 0u;
            mbar_arrive_expect_tx(bar_a, rows * Cfg::kRowBytes);
        }
        if (row_valid) {
            const uint32_t row = (uint32_t)((p.ring_off + (uint64_t)k) % p.ring_blocks);
            tma_bulk_g2s(smem_u32(xs + (size_t)t * Cfg::kRowStride), p.iq_ring + (size_t)row * BS, Cfg::kRowBytes, bar_a);
        }
        float2 Pnext = make_float2(1.0f, 0.0f);
        if (row_valid) Pnext = __ldg(phase_g[0] + k);
        mbar_wait(bar_a, tile & 1u);

        for (uint32_t ci = 0; ci < nch; ++ci) {
            const uint32_t c = c0 + ci;
            if (row_valid) {
                const float2 Pk = Pnext;
                if (ci + 1 < nch) Pnext = __ldg(phase_g[ci + 1] + k);
                const float2* __restrict__ tn = tone_s + (size_t)ci * BS;
                float2 v[BS];
#pragma unroll
                for (int m2 = 0; m2 < BS / 2; ++m2) {
                    const float4 xx = xrow[m2];
                    v[2 * m2] = cmul_unfused(make_float2(xx.x, xx.y), tn[2 * m2]);          // in[m]*tone[m]
                    v[2 * m2 + 1] = cmul_unfused(make_float2(xx.z, xx.w), tn[2 * m2 + 1]);
                }
                float2 S[8];
                ExactRows<BS>::template apply<0>(v, S);
#pragma unroll
                for (int i = 0; i < 8; ++i) Dt[i] = cmul_unfused(S[i], Pk);                   // sum*phase
                ExactRows<BS>::template apply<1>(v, S);
#pragma unroll
                for (int i = 0; i < 8; ++i) Dt[8 + i] = cmul_unfused(S[i], Pk);
                ExactRows<BS>::template apply<2>(v, S);
#pragma unroll
                for (int i = 0; i < 8; ++i) Dt[16 + i] = cmul_unfused(S[i], Pk);
                ExactRows<BS>::template apply<3>(v, S);
#pragma unroll
                for (int i = 0; i < 8; ++i) Dt[24 + i] = cmul_unfused(S[i], Pk);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) Dt[i] = make_float2(0.0f, 0.0f);
            }
            __syncthreads();

            // y[b] = sum over tap rows n ascending of D[b-31+n][n]; rows before this tile are in the carry
            float2 w = make_float2(0.0f, 0.0f);
            if (!first_tile && t < 31) w = carry_s[ci * 31 + t];
            // No predicates: a source row that lies before this tile (t - 31 + n < 0) is one of the 31 zero rows
            // in front of D, a row outside the slot (k < 0, fresh-SSBD history) was written as +0 by its thread.
            // Adding +0 is exact here: a float sum that starts from +0 (or from a carry that did) can never be
            // -0 under round-to-nearest, and x + (+0) == x for every other x. (The per-row predicates these
            // zeros replace cost 14 % of the kernel's issue slots.)
#pragma unroll
            for (int n = 0; n < 32; ++n) w = fadd2(w, D[((int)t - 31 + n) * Cfg::kDStride + n]);
            __syncwarp();
            // prefix owed to the next tile: output NT + j gets rows (NT + j - 31 + n) < NT, n ascending
            if (t < 31) {
                float2 cs = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int n = 0; n < 31; ++n) {
                    const int src = NT + t - 31 + n;  // tile-local row
                    // (src >= NT - 31 >= 31 and kt0 >= -31, so the source block is never before the slot)
                    if (src < NT) cs = fadd2(cs, D[(size_t)src * Cfg::kDStride + n]);
                }
                carry_s[ci * 31 + t] = cs;
            }

            float out = 0.0f;
            if (writes) {
                out = weaver(w.x, w.y, (uint32_t)k, p.sign[c]);
                p.audio[(size_t)c * p.af_stride + k] = out;
            }
            atomic_max_abs(p.maxbits + c, writes ? fabsf(out) : 0.0f);
            __syncthreads();
        }
    }
}

template <int BS, int NT, int CTAS, int G>
static cudaError_t launch_exact_t(const DemodLaunch& p, cudaStream_t s) {
    using Cfg = ExactCfg<BS, NT, G>;
    auto kern = demod_exact_tiled_kernel<BS, NT, CTAS, G>;
    int sms = 0;
    if (cudaError_t e = prepare_kernel((const void*)kern, (int)Cfg::kSmem, &sms); e != cudaSuccess) return e;
    const uint32_t n_out = p.b1 - p.b0;
    const uint32_t g = p.n_channels < (uint32_t)G ? p.n_channels : (uint32_t)G;
    const uint32_t groups = (p.n_channels + g - 1) / g;
    const uint32_t l = choose_tiles_per_seg(n_out, NT, groups, (uint32_t)sms * CTAS, 31);
    const uint32_t seg_out = l * NT - 31;
    dim3 grid((n_out + seg_out - 1) / seg_out, groups);
    kern<<<grid, NT, Cfg::kSmem, s>>>(p, g, l);
    return cudaGetLastError();
}

// The one-thread-per-output gather kernel above is kept as an independent second implementation
// (tests run both against the oracle); production EXACT mode uses the tiled kernel.
cudaError_t launch_demod_exact_gather(const DemodLaunch& p, cudaStream_t s) {
    if (p.b1 <= p.b0 || p.n_channels == 0) return cudaSuccess;
    dim3 grid((p.b1 - p.b0 + 127) / 128, p.n_channels);
    switch (p.block_size) {
        case 16: demod_exact_kernel<16><<<grid, 128, 0, s>>>(p); break;
        case 8: demod_exact_kernel<8><<<grid, 128, 0, s>>>(p); break;
        case 4: demod_exact_kernel<4><<<grid, 128, 0, s>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int BS, int R, int NT, int CTAS, bool PF, int G = kFastGMax>
static cudaError_t launch_fast_t(const DemodLaunch& p, uint32_t l, const FastIndirect& ind, cudaStream_t s) {
    using Cfg = FastCfg<BS, R, NT, G>;
    const uint32_t seg_out = l * Cfg::kTile - 32;
    if (p.b0 % seg_out != 0) return cudaErrorInvalidValue;  // segments sit at fixed slot-relative positions
    int sms = 0;
    if (ind.items) {
        auto kern = demod_fast_kernel<BS, R, NT, CTAS, PF, true, G>;
        if (cudaError_t e = prepare_kernel((const void*)kern, (int)Cfg::kSmem, &sms); e != cudaSuccess) return e;
        kern<<<(uint32_t)sms * CTAS, NT, Cfg::kSmem, s>>>(p, 0u, l, ind);
        return cudaGetLastError();
    }
    auto kern = demod_fast_kernel<BS, R, NT, CTAS, PF, false, G>;
    if (cudaError_t e = prepare_kernel((const void*)kern, (int)Cfg::kSmem, &sms); e != cudaSuccess) return e;
    const uint32_t n_out = p.b1 - p.b0;
    const uint32_t g = p.n_channels < (uint32_t)G ? p.n_channels : (uint32_t)G;
    const uint32_t groups = (p.n_channels + g - 1) / g;
    dim3 grid((n_out + seg_out - 1) / seg_out, groups);
    kern<<<grid, NT, Cfg::kSmem, s>>>(p, g, l, ind);
    return cudaGetLastError();
}

// Tiles per segment of the FAST-tolerance modes: a function of the slot group's size only, so that the segmentation
// (hence every rounding) of a slot does not depend on how its IQ was pushed. CWSL_TILES_PER_SEG=<n> pins it (tests
// and compute-sanitizer runs use that to force the cross-tile carry path on small inputs).
uint32_t fast_tiles_per_seg(uint32_t n_group_channels) {
    static const uint32_t forced = [] {
        const char* e = std::getenv("CWSL_TILES_PER_SEG");
        return e ? (uint32_t)std::strtoul(e, nullptr, 10) : 0u;
    }();
    if (forced > 0) return forced;
    return n_group_channels >= kSegLargeGroup ? kSegLargeGroupTiles : kSegSmallGroupTiles;
}

cudaError_t launch_demod_exact(const DemodLaunch& p, cudaStream_t s) {
    if (p.b1 <= p.b0 || p.n_channels == 0) return cudaSuccess;
    // CWSL_EXACT_SHAPE=256x2 selects 256-thread CTAs, 2 per SM (16 warps/SM); default 128x3 (12 warps/SM, cheaper
    // barriers: measured below)
    static const int shape = [] {
        const char* e = std::getenv("CWSL_EXACT_SHAPE");
        if (e && std::string(e) == "256x2") return 1;
        if (e && std::string(e) == "64x6") return 2;
        if (e && std::string(e) == "64x5") return 3;
        return 0;
    }();
    switch (p.block_size) {
        case 16:
            if (shape == 1) return launch_exact_t<16, 256, 1, 24>(p, s);
            if (shape == 2) return launch_exact_t<16, 64, 5, 16>(p, s);
            if (shape == 3) return launch_exact_t<16, 64, 4, 32>(p, s);
            return launch_exact_t<16, 128, 3, 24>(p, s);  // 69.4 KB of shared memory per CTA
        case 8: return launch_exact_t<8, 128, 3, 24>(p, s);
        case 4: return launch_exact_t<4, 128, 3, 24>(p, s);
        default: return cudaErrorInvalidValue;
    }
}

size_t fast_scratch_bytes(int device) {  // indirect launches: sms x 2 CTAs, each kFastGMax channels x kFastTile blocks
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) sms = 148;
    return (size_t)sms * 2 * kFastGMax * kFastTile * sizeof(float2);
}

cudaError_t launch_demod_fast(const DemodLaunch& p, uint32_t tiles_per_seg, const FastIndirect& ind, cudaStream_t s) {
    if (p.b1 <= p.b0 || p.n_channels == 0) return cudaSuccess;
    if (tiles_per_seg == 0) return cudaErrorInvalidValue;
    // CWSL_FAST_PREFETCH=1: keep the next block's samples/tones in registers (fetched right after the mix).
    // Measured SLOWER on B200 (563.7 vs 573.4 G ch-samples/s: the 64 extra live registers cost more than the
    // shared-memory latency they hide); kept as a documented negative result.
    static const bool pf = [] {
        const char* e = std::getenv("CWSL_FAST_PREFETCH");
        return e && e[0] == '1';
    }();
    // (<16, R=2, 128, 3 CTAs/SM, G=16> -- 12 instead of 8 warps per SM at twice the exchange traffic per block --
    // measured 506 vs 574 G ch-samples/s on B200, so the shape below stays.)
    switch (p.block_size) {
        case 16:
            return pf ? launch_fast_t<16, 4, 128, 2, true>(p, tiles_per_seg, ind, s)
                      : launch_fast_t<16, 4, 128, 2, false>(p, tiles_per_seg, ind, s);
        case 8: return launch_fast_t<8, 4, 128, 2, false>(p, tiles_per_seg, ind, s);
        case 4: return launch_fast_t<4, 4, 128, 2, false>(p, tiles_per_seg, ind, s);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------
// prepareAudio + int16 conversion (source/Instance.cpp:294-338, :238-241), all channels.
//   factor = 32767/(max+1); factor *= scale; x *= factor; q = (int16)(x + 0.5f)   (trunc toward 0)
// Each operation is a separately rounded float op, as compiled from the reference with strict
// IEEE flags. Samples past write_index are the zero tail: (int16)(0*factor + 0.5f) = 0.
// ------------------------------------------------------------------------------------------
// Each thread converts kQuantVec groups of 8 samples: the pass is pure HBM traffic (0.17 ms per 1024-channel FT8 slot
// = 6.5 TB/s alone). It runs on the receiver's post stream under the demodulation of the next receiver. 128 threads x
// 32 registers = 4096 registers per CTA is exactly what the STFT channelizer (512 x 120) leaves free on an SM, so one
// CTA of this kernel runs beside it -- and there its rate is set by the bytes those 4096 registers keep in flight,
// which is what paces the 64-receiver step (the post streams, not the channelizers, are its bottleneck). Interior CTAs
// (every group full) therefore run a branch-free path; with four groups per thread ptxas keeps four 16-byte loads in
// flight and refills as it converts (a software pipeline inside 32 registers, no spills), and a CTA lives twice as long
// before its slot has to be refilled: 43.8 instead of 44.9 ms per 64-receiver step against two groups with per-group
// predicates (three groups: 44.2; six, or an explicit rolled pipeline of depth 2 / 3: spills; 64-thread or 32-thread CTAs, 64 x 64 registers: 45.6 / 46.7 / 48.5 ms;
// tools/runs/_run53...55.sh). Alone, all shapes take 0.170-0.173 ms. (A row-walking variant with prefetch.global.L2
// stretched the channelizer by more than it saved, 0.90 ms per receiver: rejected. One cp.async.bulk.prefetch.L2 per CTA for
// the CTA 592 launches ahead -- no registers, no shared memory -- gives 43.2 instead of 43.9 ms per step but makes the pass
// alone 8 % slower, 0.188 ms: duplicate requests once every CTA of the distance is already resident; not shipped.)
#ifndef CWSL_QUANT_THREADS  // (A/B builds together with -DCWSL_CHAN_LAUNCH_REGS=...: what the channelizer leaves free)
#define CWSL_QUANT_THREADS 128
#endif
#ifndef CWSL_QUANT_VEC
#define CWSL_QUANT_VEC 4
#endif
#ifndef CWSL_QUANT_REGS
#define CWSL_QUANT_REGS 32
#endif
constexpr int kQuantThreads = CWSL_QUANT_THREADS, kQuantVec = CWSL_QUANT_VEC, kQuantRegs = CWSL_QUANT_REGS;
__device__ __forceinline__ int4 quantise8(const float4& a, const float4& b, float factor) {
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    unsigned q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] = (unsigned)(unsigned short)(short)__float2int_rz(__fadd_rn(__fmul_rn(x[e], factor), 0.5f));
    return make_int4((int)(q[0] | (q[1] << 16)), (int)(q[2] | (q[3] << 16)), (int)(q[4] | (q[5] << 16)), (int)(q[6] | (q[7] << 16)));
}

template <int kQuantVec, int kRegs>
__global__ void __maxnreg__(kRegs) quantise_kernel(QuantLaunch p) {
    const uint32_t c = blockIdx.y;
    const float maxv = __uint_as_float(p.maxbits[c]);
    float factor = __fdiv_rn(32767.0f, __fadd_rn(maxv, 1.0f));
    factor = __fmul_rn(factor, p.scale[c]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (p.factor_out) p.factor_out[c] = factor;
        if (p.max_out) p.max_out[c] = maxv;
    }
    const float* __restrict__ src = p.audio + (size_t)c * p.af_stride;
    const uint32_t pitch = p.out_pitch ? p.out_pitch : p.af_size;
    const uint32_t limit = min(pitch, p.af_size);  // packed rows end at write_index: nothing may be written behind it
    int16_t* __restrict__ dst = p.out + (size_t)c * pitch;
    const bool vec = (pitch % 8u == 0) && (p.af_stride % 4u == 0);
    const uint32_t cta0 = blockIdx.x * (kQuantThreads * 8u * kQuantVec);
    if (vec && cta0 + kQuantThreads * 8u * kQuantVec <= min(p.write_index, limit)) {
        // interior CTA (all but the last one or two of a row): every group is full -- nothing but the loads, all
        // issued before the first conversion, then the stores
        const float4* __restrict__ s4 = reinterpret_cast<const float4*>(src + cta0) + 2u * threadIdx.x;
        int4* __restrict__ d4 = reinterpret_cast<int4*>(dst + cta0) + threadIdx.x;
        float4 a[kQuantVec], b[kQuantVec];
#pragma unroll
        for (int v = 0; v < kQuantVec; ++v) {
            a[v] = __ldcs(s4 + v * (2 * kQuantThreads));      // streaming: read exactly once
            b[v] = __ldcs(s4 + v * (2 * kQuantThreads) + 1);
        }
#pragma unroll
        for (int v = 0; v < kQuantVec; ++v) __stcs(d4 + v * kQuantThreads, quantise8(a[v], b[v], factor));
        return;
    }
    // the CTAs that hold the end of the written range / of the row: one group at a time
#pragma unroll 1
    for (int v = 0; v < kQuantVec; ++v) {
        const uint32_t i0 = cta0 + v * (kQuantThreads * 8u) + threadIdx.x * 8u;
        if (i0 >= limit) continue;
        if (vec) {
            float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f), b = a;
            if (i0 + 8 <= p.write_index) {
                a = __ldcs(reinterpret_cast<const float4*>(src + i0));
                b = __ldcs(reinterpret_cast<const float4*>(src + i0 + 4));
            } else if (i0 < p.write_index) {
                float x[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = (i0 + e < p.write_index) ? src[i0 + e] : 0.0f;
                a = make_float4(x[0], x[1], x[2], x[3]);
                b = make_float4(x[4], x[5], x[6], x[7]);
            }
            __stcs(reinterpret_cast<int4*>(dst + i0), quantise8(a, b, factor));  // (0 * factor + 0.5f -> 0: the zero tail)
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t i = i0 + e;
                const float x = (i < p.write_index) ? src[i] : 0.0f;
                if (i < limit) dst[i] = (short)__float2int_rz(__fadd_rn(__fmul_rn(x, factor), 0.5f));
            }
        }
    }
}

cudaError_t launch_quantise(const QuantLaunch& p, cudaStream_t s) {
    if (p.n_channels == 0 || p.af_size == 0) return cudaSuccess;
    const uint32_t per_cta = kQuantThreads * 8u * (uint32_t)kQuantVec;
    // the zero tail behind `cover` is already zero in `out` (previous slots never wrote there): not written again
    const uint32_t cover = p.out_pitch && p.out_pitch != p.af_size ? std::min(p.out_pitch, p.af_size)
                                                                    : std::min(p.af_size, std::max(p.cover, p.write_index));
    dim3 grid(std::max(1u, (cover + per_cta - 1) / per_cta), p.n_channels);  // (>= 1: the factor / max outputs)
    quantise_kernel<kQuantVec, kQuantRegs><<<grid, kQuantThreads, 0, s>>>(p);
    return cudaGetLastError();
}

__global__ void clear_u32_kernel(unsigned* p, uint32_t n, unsigned long long* counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0u;
    if (counters && i == 0) {  // guard counters of the slot that just ended -> "last finished slot"
        counters[2] = counters[0];
        counters[3] = counters[1];
        counters[0] = counters[1] = 0ull;
    }
}
cudaError_t launch_clear_u32(unsigned* p, uint32_t n, unsigned long long* counters, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    clear_u32_kernel<<<(n + 255) / 256, 256, 0, s>>>(p, n, counters);
    return cudaGetLastError();
}

// P_c[spacing * a] of every channel's phase recurrence, [anchor][channel] (coalesced for the STFT consumers): the exact
// checkpoints from which the STFT mode's kernels replay what they need, instead of 8 bytes per audio sample of table.
__global__ void phase_anchor_kernel(const float2* __restrict__ inc, float2* __restrict__ anchors, uint32_t n_channels,
                                    uint32_t n_anchor, uint32_t spacing) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_channels) return;
    const float ir = inc[c].x, ii = inc[c].y;
    float pr = 1.0f, pi = 0.0f;
    for (uint32_t a = 0; a < n_anchor; ++a) {
        anchors[(size_t)a * n_channels + c] = make_float2(pr, pi);
        for (uint32_t j = 0; j < spacing; ++j) {  // the same operations, in the same order, as phase_table_kernel
            const float x = __fmul_rn(pr, ir), y = __fmul_rn(pi, ii);
            const float z = __fmul_rn(pr, ii), w = __fmul_rn(pi, ir);
            pr = __fsub_rn(x, y);
            pi = __fadd_rn(z, w);
        }
    }
}
cudaError_t launch_phase_anchors(const float2* phase_inc, float2* anchors, uint32_t n_channels, uint32_t n_anchor,
                                 uint32_t spacing, cudaStream_t s) {
    if (n_channels == 0 || n_anchor == 0) return cudaSuccess;
    phase_anchor_kernel<<<(n_channels + 31) / 32, 32, 0, s>>>(phase_inc, anchors, n_channels, n_anchor, spacing);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// FP32 pipe microbenchmark: register-resident FMA chains, no memory traffic.
// ------------------------------------------------------------------------------------------
template <bool PACKED>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, float a, float b, int iters) {
    // 16 independent chains acc = acc * imm + imm: the multiplier and addend are instruction immediates, like the
    // taps of the fast kernel, so an instruction reads one register (pair) only. (With three register operands
    // the scalar form measures ~46 and the packed form ~68 TFLOP/s: register-bank conflicts, not the pipe.)
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i + a, i * 0.5f + b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (PACKED) {
                    acc[i] = ffma2(acc[i], make_float2(0.999f, 0.998f), make_float2(1e-4f, 2e-4f));
                    acc[i] = ffma2(acc[i], make_float2(1.001f, 1.002f), make_float2(-1e-4f, -2e-4f));
                } else {
                    acc[i].x = fmaf(acc[i].x, 0.999f, 1e-4f);
                    acc[i].y = fmaf(acc[i].y, 0.998f, 2e-4f);
                    acc[i].x = fmaf(acc[i].x, 1.001f, -1e-4f);
                    acc[i].y = fmaf(acc[i].y, 1.002f, -2e-4f);
                }
            }
        }
    }
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) sum += acc[i].x + acc[i].y;
    if (sum == 12345.678f) out[0] = sum;  // never true; keeps the chains alive
}

cudaError_t measure_fp32_peak(float* ffma_tflops, float* ffma2_tflops) {
    int dev = 0, sms = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    float* d = nullptr;
    if ((e = cudaMalloc(&d, 4)) != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, blocks = sms * 8, threads = 256;
    const double flop = (double)blocks * threads * iters * 4 * 16 * 2 * 2;
    float best[2] = {0, 0};
    for (int variant = 0; variant < 2; ++variant) {
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            if (variant == 0)
                fp32_peak_kernel<false><<<blocks, threads>>>(d, 0.9999f, 1e-4f, iters);
            else
                fp32_peak_kernel<true><<<blocks, threads>>>(d, 0.9999f, 1e-4f, iters);
            cudaEventRecord(e1);
            if ((e = cudaEventSynchronize(e1)) != cudaSuccess) goto done;
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const float tf = (float)(flop / (ms * 1e-3) / 1e12);
            if (rep > 0 && tf > best[variant]) best[variant] = tf;
        }
    }
    if (ffma_tflops) *ffma_tflops = best[0];
    if (ffma2_tflops) *ffma2_tflops = best[1];
done:
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return e;
}

}  // namespace cwsl
