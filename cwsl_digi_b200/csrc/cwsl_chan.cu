// STFT channelizer form of the receive chain (FAST-tolerance mode CWSL_MODE_STFT, big channel groups).
//
// The reference evaluates, per channel c and output sample b (source/SSBD.hpp:160-183, gather form in
// SURVEY.md section 8 a4),
//     y_c[b] = sum_{j<512} x[16(b-31)+j] * h[j] * LO_c[16(b-31)+j],     LO_c[16k+m] = P_c[k] * tone_c[m]
// i.e. a 512-tap FIR at a 16-sample hop behind a per-channel NCO. Within one 512-sample window the NCO is a pure
// exponential to ~1e-6 (P_c[k+n] = P_c[k] e^{i n Theta_c} up to float rounding, tone_c[m] = e^{i m Theta_c/16}
// up to the rounding of its float angle), so
//     y_c[b] = P_c[b] * e^{-i 496 w_c} * W_b(w_c),     W_b(w) = sum_j x[16(b-31)+j] h[j] e^{i w j},   w_c = Theta_c/16
// and W_b is the DTFT of ONE windowed segment that all channels of the receiver share. The kernel computes it
// once per hop on a 1024-point grid (zero-padded FFT, oversampling 2) and every channel reads its own frequency
// off that grid with an 8-bin Kaiser-Bessel interpolation (the type-2 non-uniform FFT: the window is pre-divided
// by the kernel's transform). P_c[b] is still the reference's own drifting float phase recurrence (phase table),
// so the NCO drift (|P| = 1.0004 after one FT8 slot) is reproduced; only the *within-window* deviation from a pure
// exponential is dropped. Measured against the reference chain (tools/chan_proto.py, tests): residual <= -120 dB
// with a kernel of width 7, <= 1 int16 LSB. Per channel-sample this costs ~1.5 FMA-pipe instructions instead of 24 (folded
// direct form), which moves the path from the FP32 pipe to shared-memory/HBM bandwidth.
//
// (Written for 192 kHz: window 512, hop 16, grid 1024. At 96 / 48 kHz the window is 256 / 128 taps, the hop 8 / 4
// samples and the grid 512 / 256 bins; see ChanGeo.)
//
// Kernel: persistent and warp-specialised, one 512-thread CTA per SM owning a contiguous run of batches of hops.
//   warps 0-7  (producers) transform the hops of a batch: the windowed samples straight from the IQ ring (zero
//              history before the slot start, source/Instance.cpp:251), N-point FFT as (N/32) x 32 -- two
//              register-resident radix-2 DIF passes with one transpose through the hop's own spectrum buffer --
//              window, inter-pass twiddles and the centring rotation i^q from small L1-resident tables;
//   warps 8-15 (consumers) own one work ITEM per thread for the whole launch: up to 4 channels that are neighbours on
//              the FFT grid (cwsl_tables.hpp chan_items) and therefore read ONE 12-bin window -- 6 x LDS.128 per hop
//              serve four channels (each channel alone would need 4: shared-memory wavefronts were this kernel's
//              limit). Weights (4 x 9, zero outside the kernel's support), rot*P and phase_inc live in registers;
//              per hop and channel 9 x FFMA2, Weaver select (source/SSBD.hpp:132-135), R <- R * phase_inc with R
//              re-read from the exact phase table every 128 hops, one 32-byte store per channel and 8 hops, max|x| in
//              a register until the end (one atomicMax per channel and CTA);
//   registers move from the FFT warps to the interpolation warps at the role split (setmaxnreg);
//   spectra are handed over through a ring of three shared-memory buffers with named barriers (bar.arrive /
//   bar.sync), so the FFT warps run up to two batches ahead of the interpolation warps.
#include "cwsl_kernels.hpp"
#include "cwsl_ptx.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <set>
#include <utility>

namespace cwsl {

namespace {

constexpr int kRow = 34;  // float2 per transpose row: conflict-free 64-bit column writes AND 128-bit row reads

// Geometry per receiver rate: window L = FiltOrder = 32 * BS taps, grid N = 2 L, FFT as A x 32 with A = N / 32.
// One warp transforms HW = 32 / A hops at a time (pass 1: every lane runs HW A-point DFTs; pass 2: lane -> (hop,
// q1), one 32-point DFT each), so all 32 lanes are busy in both passes at every rate.
template <int BS>
struct ChanGeo {
    static constexpr int kL = 32 * BS;       // 512 / 256 / 128
    static constexpr int kN = 2 * kL;        // 1024 / 512 / 256
    static constexpr int kA = kN / 32;       // 32 / 16 / 8
    static constexpr int kHW = 32 / kA;      // hops per warp and batch: 1 / 2 / 4
    static constexpr int kHB = 8 * kHW;      // hops per batch (8 FFT warps): 8 / 16 / 32
    // float2 per hop buffer: A rows of 34 during the transpose, N (+8 wrapped) bins after; the 8-row geometry gets 8
    // more so that the two hops of a half-warp sit 16 banks apart
    static constexpr int kHop = kA * kRow + (kA == 8 ? 8 : 0);
    static constexpr int kWrap = kChanItemBins - 2;  // bins 0..9 again behind bin N-1: a 12-bin window never wraps
    static_assert(kN + kWrap <= kHop, "wrapped stencil bins must fit");
};

// d * (wr + i wi) as two packed instructions: FMUL2 with the broadcast real part, then FFMA2 of the rotated operand
// (-d.y, d.x) -- ptxas folds the swap/negate into the FFMA2 operand modifiers -- with the broadcast imaginary part.
__device__ __forceinline__ float2 cmul_pk(float2 d, float wr, float wi) {
    return ffma2(make_float2(-d.y, d.x), bc(wi), fmul2(d, bc(wr)));
}

// W32^k = exp(-2 pi i k / 32), k = 0..15
__device__ __forceinline__ float2 mul_w32(float2 d, int k) {
    constexpr float c1 = 0.98078528040323044913f, s1 = 0.19509032201612826785f;
    constexpr float c2 = 0.92387953251128675613f, s2 = 0.38268343236508977173f;
    constexpr float c3 = 0.83146961230254523708f, s3 = 0.55557023301960222474f;
    constexpr float c4 = 0.70710678118654752440f;
    switch (k) {
        case 0: return d;
        case 1: return cmul_pk(d, c1, -s1);
        case 2: return cmul_pk(d, c2, -s2);
        case 3: return cmul_pk(d, c3, -s3);
        case 4: return cmul_pk(d, c4, -c4);
        case 5: return cmul_pk(d, s3, -c3);
        case 6: return cmul_pk(d, s2, -c2);
        case 7: return cmul_pk(d, s1, -c1);
        case 8: return make_float2(d.y, -d.x);
        case 9: return cmul_pk(d, -s1, -c1);
        case 10: return cmul_pk(d, -s2, -c2);
        case 11: return cmul_pk(d, -s3, -c3);
        case 12: return cmul_pk(d, -c4, -c4);
        case 13: return cmul_pk(d, -c3, -s3);
        case 14: return cmul_pk(d, -c2, -s2);
        default: return cmul_pk(d, -c1, -s1);
    }
}


template <int A>
__host__ __device__ constexpr int bitrev(int i) {  // over log2(A) bits
    int r = 0;
    for (int b = 1, m = A >> 1; b < A; b <<= 1, m >>= 1)
        if (i & b) r |= m;
    return r;
}

// In-register A-point forward DFT (A = 8, 16, 32) on v[OFF .. OFF+A), radix-2 decimation in frequency;
// X[bitrev<A>(i)] is left in v[OFF + i]. UPPER_ZERO: v[OFF+A/2 .. OFF+A) are known zeros on entry (the zero-padded
// half of the window). W_A^k = W32^(k * 32/A).
template <int A, int OFF, bool UPPER_ZERO>
__device__ __forceinline__ void fft_dif(float2 (&v)[32]) {
    constexpr int H = A / 2;
    if constexpr (UPPER_ZERO) {
#pragma unroll
        for (int k = 0; k < H; ++k) v[OFF + k + H] = mul_w32(v[OFF + k], k * (16 / H));
    } else {
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const float2 a = v[OFF + k], b = v[OFF + k + H];
            v[OFF + k] = fadd2(a, b);
            v[OFF + k + H] = mul_w32(fsub2(a, b), k * (16 / H));
        }
    }
#pragma unroll
    for (int half = H / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int g = 0; g < A; g += 2 * half) {
#pragma unroll
            for (int k = 0; k < half; ++k) {
                const float2 a = v[OFF + g + k], b = v[OFF + g + k + half];
                v[OFF + g + k] = fadd2(a, b);
                v[OFF + g + k + half] = mul_w32(fsub2(a, b), k * (16 / half));
            }
        }
    }
}
template <int A, bool UPPER_ZERO>
__device__ __forceinline__ void fft_dif_all(float2 (&v)[32]) {  // the 32/A transforms a lane holds in pass 1
    fft_dif<A, 0, UPPER_ZERO>(v);
    if constexpr (A <= 16) fft_dif<A, A, UPPER_ZERO>(v);
    if constexpr (A <= 8) {
        fft_dif<A, 2 * A, UPPER_ZERO>(v);
        fft_dif<A, 3 * A, UPPER_ZERO>(v);
    }
}

constexpr int kFftThreadsFwd = 256, kIntThreadsFwd = 256;
// named barriers (0 is __syncthreads): spectrum buffer s is FULL / EMPTY
constexpr int kBufs = 3;          // spectrum buffers in flight between the FFT warps and the interpolation warps
// Two hand-over groups: FFT warps 0-3 transform the first half of a batch's hops, warps 4-7 the second half, each half
// with its own FULL / EMPTY barrier per spectrum buffer (ids 1..6 and 7..12). The interpolation warps start on a half
// as soon as ITS four warps are done and hand it back as soon as they have read it, instead of waiting for -- and
// releasing -- all eight at once: 2 % at both ends of the channel range (0.584 -> 0.573 ms at 1024 channels, 0.339 ->
// 0.331 ms at 64). (Starting the second group half a hop late so that the two FFT warps of a scheduler alternate
// between arithmetic and transposes made no difference: the skew does not survive.)
constexpr int kGroupThreads = kFftThreadsFwd / 2 + kIntThreadsFwd;
__device__ __forceinline__ int bar_full2(uint32_t s, int grp) { return 1 + 2 * (int)s + grp; }
__device__ __forceinline__ int bar_empty2(uint32_t s, int grp) { return 1 + 2 * kBufs + 2 * (int)s + grp; }
__device__ __forceinline__ void bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ void stg256(float* dst, const float (&o)[8]) {  // one 32-byte sector per lane (STG.256)
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]),
                 "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                 : "memory");
}

constexpr int kFftWarps = 8;
constexpr int kFftThreads = 32 * kFftWarps;
constexpr int kIntThreads = 256;  // interpolation threads per CTA, each owns one work item (<= 4 channels) for the whole launch
constexpr int kThreads = kFftThreads + kIntThreads;
// Registers: the kernel starts with kLaunchRegs per thread; at the role split the FFT warps give registers back and the
// interpolation warps (4 x 9 weights + a 12-bin window + 32 outputs in flight per thread) take them (setmaxnreg).
// 512 x 120 leaves 4 K registers per SM for a CTA of the (HBM-bound) quantise kernel of the previous receiver, which
// then runs underneath this (shared-memory-bound) kernel instead of after it.
#ifndef CWSL_CHAN_LAUNCH_REGS  // (A/B builds: -DCWSL_CHAN_LAUNCH_REGS= -DCWSL_CHAN_FFT_REGS= -DCWSL_CHAN_INT_REGS=)
#define CWSL_CHAN_LAUNCH_REGS 120
#define CWSL_CHAN_FFT_REGS 88
#define CWSL_CHAN_INT_REGS 152
#endif
constexpr int kLaunchRegs = CWSL_CHAN_LAUNCH_REGS, kFftRegs = CWSL_CHAN_FFT_REGS, kIntRegs = CWSL_CHAN_INT_REGS;
static_assert(kFftThreads * kFftRegs + kIntThreads * kIntRegs <= kThreads * kLaunchRegs, "register budget");
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
// IQ staging ring in shared memory: stages of one batch worth of new samples (HB blocks = 1 KB at every rate),
// brought in by the TMA bulk-copy engine iq_prefetch() batches ahead of the FFT warps. 16 stages at 192 kHz: a batch reads the
// newest stage and the <= 4 before it (the 31 blocks of filter history), 3 are in flight, and the FFT warps are never
// more than kBufs batches apart (they meet the interpolation warps at the spectrum ring), so a stage is overwritten
// long after its last reader -- no "empty" barrier is needed.
template <int BS>
__host__ __device__ constexpr int iq_stages() { return BS == 4 ? 8 : 16; }    // (48 kHz: batches of 32 hops; the smaller ring makes room
template <int BS>
__host__ __device__ constexpr int iq_prefetch() { return BS == 4 ? 1 : 3; }  //  for the 210 KB of spectrum buffers of that geometry)
template <int BS>
__host__ __device__ constexpr size_t spec_bytes() {  // 3 buffers x (8 FFT warps x HW hops) x hop buffer: 204 KB at every rate
    return (size_t)kBufs * ChanGeo<BS>::kHB * ChanGeo<BS>::kHop * 8;
}
template <int BS>
__host__ __device__ constexpr size_t iq_stage_bytes() { return (size_t)ChanGeo<BS>::kHB * BS * 8; }  // 1 KB
// Per-lane constants of the FFT warps, one padded record per lane (conflict-free LDS.128): the inter-pass twiddle
// W_N^(lane q1) i^q1 split as T1[q1 >> 2] * T2[q1 & 3] (A/4 + 4 complex values instead of A, so they are re-read from
// shared memory per hop instead of from L1 per use), then the lane's A/2 window taps.
template <int BS>
__host__ __device__ constexpr int lane_rec_floats() { return ChanGeo<BS>::kA + 8; }
template <int BS>
__host__ __device__ constexpr int lane_rec_stride() { return ChanGeo<BS>::kA + 12; }  // floats; 44 / 28 / 20: odd multiples of 4
template <int BS>
__host__ __device__ constexpr size_t chan_smem_bytes() {
    return spec_bytes<BS>() + iq_stages<BS>() * iq_stage_bytes<BS>() + iq_stages<BS>() * 8 + 32 * lane_rec_stride<BS>() * 4;
}

// Complex multiply with a FIXED contraction pattern: the per-hop NCO step R <- R * phase_inc between anchors must
// give the same bits wherever it is evaluated (in the hop loop, or replayed from the anchor at the start of a CTA's
// run), or the output would depend on how a slot is cut into launches.
__device__ __forceinline__ float2 cmul_fix(float2 a, float2 b) {
    return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)), __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}

// Persistent, warp-specialised: CTA j owns a contiguous run of batches (kHB hops each). Warps 0..7 (producers)
// compute the spectra of the batch's hops into spectrum buffer s; warps 8..15 (consumers) read every channel's
// frequency off those spectra while the producers are already transforming the next batches. Hand-over by named
// barriers (bar.arrive / bar.sync), no __syncthreads in the loop. The IQ samples reach the FFT warps through a
// shared-memory ring filled by cp.async.bulk (TMA) with mbarrier completion: each sample crosses L2 -> SM once per
// CTA run instead of once per hop that covers it (32x).
template <int BS>
__global__ void __maxnreg__(kLaunchRegs)
    demod_chan_kernel(DemodLaunch p, ChanLaunch c, uint32_t n_batches, uint32_t batches_per_cta) {
    using G = ChanGeo<BS>;
    constexpr int A = G::kA, HW = G::kHW, HB = G::kHB, HOP = G::kHop, N = G::kN;
    constexpr int P0 = (31 + HB - 1) / HB;   // stages of filter history in front of a batch's own stage: 4 / 2 / 1
    constexpr int kIqStages = iq_stages<BS>(), kIqPrefetch = iq_prefetch<BS>();
    constexpr uint32_t RB = kIqStages * HB;  // SSBD blocks in the shared-memory IQ ring (a power of two)
    constexpr uint32_t kStageBytes = (uint32_t)iq_stage_bytes<BS>();
    static_assert((RB & (RB - 1)) == 0 && P0 * HB >= 31 && P0 + 1 + kIqPrefetch + kBufs + 2 <= kIqStages, "IQ ring geometry");
    extern __shared__ __align__(128) unsigned char smem[];
    float2* spec = reinterpret_cast<float2*>(smem);
    float2* iq_s = reinterpret_cast<float2*>(smem + spec_bytes<BS>());
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + spec_bytes<BS>() + kIqStages * iq_stage_bytes<BS>());
    float* lane_rec = reinterpret_cast<float*>(smem + spec_bytes<BS>() + kIqStages * iq_stage_bytes<BS>() + kIqStages * 8);
    constexpr int REC = lane_rec_floats<BS>(), RSTR = lane_rec_stride<BS>();
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t s0 = blockIdx.x * batches_per_cta;
    const uint32_t s1 = min(n_batches, s0 + batches_per_cta);
    if (s0 >= s1) return;
    if (t == 0) {
#pragma unroll
        for (int i = 0; i < kIqStages; ++i) mbar_init(smem_u32(bars + i), 1);
        fence_mbar_init();
    }
    for (uint32_t i = t; i < 32u * REC; i += kThreads) {  // lane records: T1[A/4], T2[4] (float2 each), window[A/2]
        const uint32_t ln = i / REC, f = i % REC;
        float val;
        if (f < (uint32_t)(A / 2)) {
            const float2 w = __ldg(c.twiddle + 32 * (4 * (f >> 1)) + ln);   // q1 = 4a: W_N^(lane 4a), i^(4a) = 1
            val = (f & 1u) ? w.y : w.x;
        } else if (f < (uint32_t)(A / 2 + 8)) {
            const float2 w = __ldg(c.twiddle + 32 * ((f - A / 2) >> 1) + ln);  // q1 = b < 4: W_N^(lane b) i^b
            val = (f & 1u) ? w.y : w.x;
        } else {
            val = __ldg(c.window + 32 * (f - A / 2 - 8) + ln);
        }
        lane_rec[ln * RSTR + f] = val;
    }
    __syncthreads();

    if (warp < (uint32_t)kFftWarps) {
        // ================= producers: warp w transforms hops w*HW .. w*HW+HW-1 of every batch =================
        setmaxnreg_dec<kFftRegs>();
        const uint32_t nb = s1 - s0;
        // local stage l of this CTA holds the slot-relative blocks kbase0 + HB*l .. +HB-1; batch `it` (its own new
        // samples are stage P0 + it) reads stages it .. it + P0
        const long long kbase0 = (long long)p.b0 + ((long long)s0 - P0) * HB;
        // Stages are issued strictly in order, so the ring row of the next one is kept incrementally (one 64-bit
        // modulo per launch instead of one per stage: the issuing lane's ~100-instruction division held up the whole
        // of warp 0 -- and with it, through the batch barrier, all FFT warps -- for ~600 cycles per batch).
        uint32_t next_row;
        {
            long long m = ((long long)p.ring_off + kbase0) % (long long)p.ring_blocks;
            if (m < 0) m += p.ring_blocks;  // (blocks before the slot: whatever the ring holds, masked below)
            next_row = (uint32_t)m;
        }
        auto issue_stage = [&](uint32_t l) {
            const uint32_t row = next_row, slot = l % kIqStages;
            next_row += (uint32_t)HB;
            if (next_row >= p.ring_blocks) next_row -= p.ring_blocks;  // (HB <= 32 < 64 <= ring_blocks)
            const uint32_t dst = smem_u32(iq_s + (size_t)slot * HB * BS), bar_a = smem_u32(bars + slot);
            mbar_arrive_expect_tx(bar_a, kStageBytes);
            const uint32_t first = min((uint32_t)HB, p.ring_blocks - row);  // the IQ ring may wrap inside the stage
            tma_bulk_g2s(dst, p.iq_ring + (size_t)row * BS, first * BS * 8u, bar_a);
            if (first < (uint32_t)HB) tma_bulk_g2s(dst + first * BS * 8u, p.iq_ring, (HB - first) * BS * 8u, bar_a);
        };
        if (t == 0)
            for (uint32_t l = 0; l <= (uint32_t)P0 + min((uint32_t)kIqPrefetch, nb - 1); ++l) issue_stage(l);
        // window of hop b = IQ blocks b-31 .. b (BS samples each); element j = 32 j1 + lane of the window sits in
        // block b-31 + j1*(32/BS) + lane/BS, sample lane % BS
        long long blk = (long long)p.b0 + (long long)s0 * HB + (long long)warp * HW - 31 + (long long)(lane / BS);
        uint32_t lb = (uint32_t)P0 * HB + warp * HW - 31 + lane / BS;  // the same block, counted from kbase0
        const float4* rec4 = reinterpret_cast<const float4*>(lane_rec + lane * RSTR);
        uint32_t s = 0, waited = 0;
        for (uint32_t it = 0; it < nb; ++it, s = (s + 1 == (uint32_t)kBufs) ? 0u : s + 1) {
            if (t == 0 && it > 0 && it + kIqPrefetch < nb) issue_stage((uint32_t)P0 + it + kIqPrefetch);
            while (waited <= (uint32_t)P0 + it) {  // stages arrive in order; each is waited for once
                mbar_wait(smem_u32(bars + (waited % kIqStages)), (waited / kIqStages) & 1u);
                ++waited;
            }
            float2* wbuf = spec + ((size_t)s * HB + (size_t)warp * HW) * HOP;  // this warp's HW hop buffers
            float2 v[32];
            float wv[A / 2];  // this lane's window taps
#pragma unroll
            for (int j = 0; j < A / 8; ++j) {
                const float4 f = rec4[(A / 2 + 8) / 4 + j];
                wv[4 * j] = f.x, wv[4 * j + 1] = f.y, wv[4 * j + 2] = f.z, wv[4 * j + 3] = f.w;
            }
            // pass 1: lane = j2; per hop an A-point DFT over j1 of u[32 j1 + j2], u = x * window (j1 >= A/2: zero padding)
            // Common case (warp-uniform): the hop's 32 blocks neither wrap in the shared-memory ring nor reach back
            // before the slot -> one base pointer, compile-time offsets, no masks.
            const uint32_t lbw = (lb - lane / BS) & (RB - 1u);  // the warp's first block of this batch in the ring
            if (lbw + 31u + (uint32_t)HW <= RB && blk - (long long)(lane / BS) >= 0) {
                const float2* src = iq_s + (size_t)(lbw + lane / BS) * BS + (lane % BS);
#pragma unroll
                for (int sub = 0; sub < HW; ++sub)
#pragma unroll
                    for (int j1 = 0; j1 < A / 2; ++j1)
                        v[sub * A + j1] = fmul2(src[(sub + j1 * (32 / BS)) * BS], bc(wv[j1]));
            } else {
#pragma unroll
                for (int sub = 0; sub < HW; ++sub) {
#pragma unroll
                    for (int j1 = 0; j1 < A / 2; ++j1) {
                        const int boff = sub + j1 * (32 / BS);     // block offset of this element from the lane's base block
                        float2 x = iq_s[(size_t)((lb + boff) & (RB - 1u)) * BS + (lane % BS)];
                        if (blk + boff < 0) x = make_float2(0.0f, 0.0f);  // zero history before the slot (fresh SSBD)
                        v[sub * A + j1] = fmul2(x, bc(wv[j1]));
                    }
                }
            }
            fft_dif_all<A, true>(v);
            if (it >= (uint32_t)kBufs) bar_sync(bar_empty2(s, (int)(warp >> 2)), kGroupThreads);
            // twiddle W_N^(j2 q1) * i^q1 = T1[q1 >> 2] * T2[q1 & 3] and transpose through the hop buffer (rows of 34:
            // conflict-free both ways)
            {
                float2 t1[A / 4], t2[4];
#pragma unroll
                for (int j = 0; j < A / 8; ++j) {
                    const float4 f = rec4[j];
                    t1[2 * j] = make_float2(f.x, f.y), t1[2 * j + 1] = make_float2(f.z, f.w);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 f = rec4[A / 8 + j];
                    t2[2 * j] = make_float2(f.x, f.y), t2[2 * j + 1] = make_float2(f.z, f.w);
                }
#pragma unroll
                for (int sub = 0; sub < HW; ++sub) {
#pragma unroll
                    for (int k = 0; k < A; ++k) {
                        const int q1 = bitrev<A>(k);
                        float2 a = v[sub * A + k];
                        if (q1 >> 2) a = cmul_pk(a, t1[q1 >> 2].x, t1[q1 >> 2].y);
                        if (q1 & 3) a = cmul_pk(a, t2[q1 & 3].x, t2[q1 & 3].y);
                        wbuf[sub * HOP + q1 * kRow + lane] = a;
                    }
                }
            }
            __syncwarp();
            // pass 2: lane -> (hop sub2, q1); 32-point DFT over j2 -> bins q1 + A q2 of that hop
            const uint32_t sub2 = lane / A, q1 = lane % A;
            float2* buf = wbuf + sub2 * HOP;
#pragma unroll
            for (int j2 = 0; j2 < 32; j2 += 2) {
                const float4 two = *reinterpret_cast<const float4*>(buf + q1 * kRow + j2);
                v[j2] = make_float2(two.x, two.y);
                v[j2 + 1] = make_float2(two.z, two.w);
            }
            __syncwarp();
            fft_dif<32, 0, false>(v);
#pragma unroll
            for (int k = 0; k < 32; ++k) buf[A * bitrev<32>(k) + q1] = v[k];
            // bins 0..9 again behind bin N-1, so the 12-bin windows (even start) never wrap: bin q1 + A*q2
            if constexpr (A >= G::kWrap) {
                if (q1 < (uint32_t)G::kWrap) buf[N + q1] = v[0];
            } else {
                buf[N + q1] = v[0];                                                   // q2 = 0
                if (q1 < (uint32_t)(G::kWrap - A)) buf[N + A + q1] = v[bitrev<32>(1)];  // q2 = 1
            }
            bar_arrive(bar_full2(s, (int)(warp >> 2)), kGroupThreads);
            blk += HB;
            lb += HB;
        }
    } else {
        // ================= consumers: thread tid owns work item tid (<= 4 neighbouring channels) =================
        setmaxnreg_inc<kIntRegs>();
        constexpr int M = kChanItemMembers, T = kChanItemTaps;
        constexpr int kShift[M] = {0, 0, 1, 2};
        const uint32_t tid = t - kFftThreads;
        const bool act = tid < c.n_items;
        const ChanItem* item = c.items + (act ? tid : 0u);
        float w[M][T];
        float2 pinc[M], R[M];
        float sgn[M], mx[M];
        uint32_t chn[M], ea[M];  // ea (guard): integer sum of the scaled octet energies of the current segment
        uint32_t boff;
        {
            const float4* q = reinterpret_cast<const float4*>(item);
            const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4);
            boff = (uint32_t)__float_as_int(q0.x);
            chn[0] = (uint32_t)__float_as_int(q1.x), chn[1] = (uint32_t)__float_as_int(q1.y);
            chn[2] = (uint32_t)__float_as_int(q1.z), chn[3] = (uint32_t)__float_as_int(q1.w);
            sgn[0] = q2.x, sgn[1] = q2.y, sgn[2] = q2.z, sgn[3] = q2.w;
            pinc[0] = make_float2(q3.x, q3.y), pinc[1] = make_float2(q3.z, q3.w);
            pinc[2] = make_float2(q4.x, q4.y), pinc[3] = make_float2(q4.z, q4.w);
            const float* wp = reinterpret_cast<const float*>(item) + 28;  // w[][] starts at byte 112
#pragma unroll
            for (int j = 0; j < M; ++j)
#pragma unroll
                for (int i = 0; i < T; ++i) w[j][i] = __ldg(wp + j * T + i);
#pragma unroll
            for (int j = 0; j < M; ++j) {
                if (!act) chn[j] = 0xffffffffu;
                R[j] = make_float2(0.0f, 0.0f);
                mx[j] = 0.0f;
                ea[j] = 0u;
            }
        }
        // R = P_c[128 a] * rot_c from the exact recurrence (anchor table), then R <- R * phase_inc per hop
        auto anchor = [&](int j, uint32_t a) {
            const float2 P = __ldg(c.anchors + (size_t)a * c.anchor_stride + chn[j]);
            const float2 rot = __ldg(reinterpret_cast<const float2*>(item) + 10 + j);  // rot[][] starts at byte 80
            return cmul_fix(P, rot);
        };
        const uint32_t bb0 = p.b0 + s0 * HB;
        {   // a run that starts between two anchors replays the steps from the anchor before it: same bits as a run
            // that came through them
            const uint32_t a0 = bb0 / kChanAnchorHops, n_replay = bb0 % kChanAnchorHops;
#pragma unroll
            for (int j = 0; j < M; ++j) {
                if (chn[j] == 0xffffffffu) continue;
                float2 r = anchor(j, a0);
#pragma unroll 1
                for (uint32_t i = 0; i < n_replay; ++i) r = cmul_fix(r, pinc[j]);
                R[j] = r;
            }
        }
        // dynamic-range guard statistics (cwsl_guard.cu): per segment max|y| and an integer energy sum
        const bool guard = c.seg_blocks != 0u;
        uint32_t cur_seg = guard ? (bb0 - p.b0) / c.seg_blocks : 0u;
        uint32_t next_seg_b = guard ? p.b0 + (cur_seg + 1u) * c.seg_blocks : 0xffffffffu;
        float s2 = guard ? __ldg(c.seg_scale + cur_seg) : 0.0f;
        auto flush = [&]() {
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const uint32_t ch = chn[j];
                if (ch == 0xffffffffu) continue;
                const unsigned bits = __float_as_uint(mx[j]);
                if (guard) {
                    const size_t i = (size_t)cur_seg * c.stat_stride + ch;
                    if (bits != 0u) atomicMax(c.seg_max + i, bits);
                    if (ea[j] != 0u) atomicAdd(c.seg_energy + i, ea[j]);
                } else if (bits != 0u && bits > __ldcg(p.maxbits + ch)) {
                    atomicMax(p.maxbits + ch, bits);
                }
                mx[j] = 0.0f;
                ea[j] = 0u;
            }
        };
        uint32_t s = 0;
        for (uint32_t i = s0; i < s1; ++i, s = (s + 1 == (uint32_t)kBufs) ? 0u : s + 1) {
            const uint32_t bb = p.b0 + i * HB;
            if (i != s0 && (bb % kChanAnchorHops) == 0u) {
#pragma unroll
                for (int j = 0; j < M; ++j)
                    if (chn[j] != 0xffffffffu) R[j] = anchor(j, bb / kChanAnchorHops);
            }
            if (bb >= next_seg_b) {  // (segments are multiples of 32 hops: a batch never straddles two)
                flush();
                ++cur_seg;
                next_seg_b += c.seg_blocks;
                s2 = __ldg(c.seg_scale + cur_seg);
            }
            {
                float out[M][8];
                auto do_hop = [&](int h, const unsigned char* base) {
                        const float4* bins = reinterpret_cast<const float4*>(base + (size_t)h * HOP * 8);
                        float2 bn[kChanItemBins];
#pragma unroll
                        for (int q = 0; q < kChanItemBins / 2; ++q) {
                            const float4 two = bins[q];
                            bn[2 * q] = make_float2(two.x, two.y);
                            bn[2 * q + 1] = make_float2(two.z, two.w);
                        }
#pragma unroll
                        for (int j = 0; j < M; ++j) {
                            float2 acc = fmul2(bn[kShift[j]], bc(w[j][0]));
#pragma unroll
                            for (int q = 1; q < T; ++q) acc = ffma2(bn[kShift[j] + q], bc(w[j][q]), acc);
                            const float2 r = R[j];
                            // audio[b] = {+Re, -Im*sign, -Re, +Im*sign}[b & 3] of acc*R; b8 is a multiple of 4
                            float o;
                            if ((h & 3) == 0) o = __fmaf_rn(acc.x, r.x, -__fmul_rn(acc.y, r.y));
                            else if ((h & 3) == 1) o = -__fmaf_rn(acc.x, r.y, __fmul_rn(acc.y, r.x)) * sgn[j];  // (x * +-1: exact)
                            else if ((h & 3) == 2) o = -__fmaf_rn(acc.x, r.x, -__fmul_rn(acc.y, r.y));
                            else o = __fmaf_rn(acc.x, r.y, __fmul_rn(acc.y, r.x)) * sgn[j];
                            out[j][h] = o;
                            R[j] = cmul_fix(r, pinc[j]);
                        }
                };
                auto store_octet = [&](uint32_t b8) {
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        if (chn[j] == 0xffffffffu) continue;
                        float* dst = p.audio + (size_t)chn[j] * p.af_stride + b8;
                        float e8 = 0.0f;
                        if (b8 + 8 <= p.b1) {  // whole octet, 32-byte aligned row segment
                            stg256(dst, out[j]);
                            float m = mx[j];
#pragma unroll
                            for (int h = 0; h < 8; ++h) {
                                m = fmaxf(m, fabsf(out[j][h]));
                                e8 = __fmaf_rn(out[j][h], out[j][h], e8);
                            }
                            mx[j] = m;
                        } else if (b8 < p.b1) {  // b1 is a multiple of 4: the first quartet only
                            *reinterpret_cast<float4*>(dst) = make_float4(out[j][0], out[j][1], out[j][2], out[j][3]);
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                mx[j] = fmaxf(mx[j], fabsf(out[j][h]));
                                e8 = __fmaf_rn(out[j][h], out[j][h], e8);
                            }
                        }
                        // scaled so that the guard's keep-threshold is 128 per hop on average; the integer sum is exact,
                        // hence independent of how the segment is spread over CTAs and launches
                        if (guard) ea[j] += __float2uint_rz(fminf(__fmul_rn(e8, s2), 1048576.0f));
                    }
                };
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    bar_sync(bar_full2(s, half), kGroupThreads);
                    if (act) {
                        if constexpr (HB == 8) {  // half a batch = half an octet
                            const unsigned char* base = smem + (size_t)s * HB * HOP * 8 + boff;
#pragma unroll
                            for (int h = 0; h < 4; ++h) do_hop(4 * half + h, base);
                            if (half == 1) store_octet(bb);
                        } else {
#pragma unroll 1
                            for (int oct = half * (HB / 16); oct < (half + 1) * (HB / 16); ++oct) {
                                const unsigned char* base = smem + ((size_t)s * HB + 8 * oct) * HOP * 8 + boff;
#pragma unroll
                                for (int h = 0; h < 8; ++h) do_hop(h, base);
                                store_octet(bb + 8 * oct);
                            }
                        }
                    }
                    if (i + kBufs < s1) bar_arrive(bar_empty2(s, half), kGroupThreads);  // (nobody waits for the last ones)
                }
            }
        }
        flush();
    }
}

std::mutex g_attr_mu;
std::set<std::pair<int, const void*>> g_attr_done;

cudaError_t prepare(const void* kern, size_t smem_bytes, int* sms) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_attr_mu);
    if (g_attr_done.count({dev, kern})) return cudaSuccess;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100)) != cudaSuccess) return e;
    g_attr_done.insert({dev, kern});
    return cudaSuccess;
}

template <int BS>
cudaError_t launch_bs(const DemodLaunch& p, const ChanLaunch& c, cudaStream_t s) {
    constexpr int HB = ChanGeo<BS>::kHB;
    auto kern = demod_chan_kernel<BS>;
    int sms = 0;
    cudaError_t e = prepare(reinterpret_cast<const void*>(kern), chan_smem_bytes<BS>(), &sms);
    if (e != cudaSuccess) return e;
    const uint32_t n_batches = (p.b1 - p.b0 + HB - 1) / HB;
    // one CTA per SM, contiguous runs of batches (anchored phase recurrence, sequential IQ reads)
    const uint32_t per_cta = (n_batches + (uint32_t)sms - 1) / (uint32_t)sms;
    const uint32_t grid = (n_batches + per_cta - 1) / per_cta;
    kern<<<grid, kThreads, chan_smem_bytes<BS>(), s>>>(p, c, n_batches, per_cta);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_demod_chan(const DemodLaunch& p, const ChanLaunch& c, cudaStream_t s) {
    if (p.n_channels == 0 || c.n_items == 0 || c.n_items > kChanMaxItems || !c.items || p.ring_blocks < 64 ||
        (p.b0 & 31u) || (p.b1 & 3u) || (p.af_stride & 7u) || !c.anchors ||
        (c.seg_blocks != 0u && (p.b0 % c.seg_blocks != 0u || (c.seg_blocks & 31u) || !c.seg_max || !c.seg_energy || !c.seg_scale)))
        return cudaErrorInvalidValue;
    if (p.b1 <= p.b0) return cudaSuccess;
    switch (p.block_size) {
        case 16: return launch_bs<16>(p, c, s);
        case 8: return launch_bs<8>(p, c, s);
        case 4: return launch_bs<4>(p, c, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace cwsl
