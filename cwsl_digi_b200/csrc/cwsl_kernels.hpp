// Launch interface of the sm_100a demodulator kernels (cwsl_kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace cwsl {

// Everything one demodulator launch needs. All pointers are device pointers.
struct DemodLaunch {
    // IQ ring: ring_blocks SSBD blocks of block_size complex samples each (float2 = I,Q).
    const float2* iq_ring = nullptr;
    uint32_t ring_blocks = 0;
    uint32_t ring_off = 0;    // ring row of the slot's block 0
    uint32_t block_size = 0;  // 16 @192 kHz, 8 @96 kHz, 4 @48 kHz
    // Slot-relative SSBD block range [b0, b1) to demodulate (= audio sample range); multiples of 4.
    // Blocks < 0 are the fresh-SSBD zero history (source/Instance.cpp:251).
    uint32_t b0 = 0, b1 = 0;
    // Channels of the slot group.
    uint32_t n_channels = 0;
    const float2* tone = nullptr;            // [n_channels][block_size]       exact mode
    const float2* const* phase = nullptr;    // [n_channels] -> phase table (>= b1 entries)
    const float* sign = nullptr;             // [n_channels] +1 USB / -1 LSB
    float* audio = nullptr;                  // [n_channels][af_stride] float audio (pre-normalise)
    size_t af_stride = 0;
    unsigned* maxbits = nullptr;             // [n_channels] bit pattern of max|x| so far
};

struct QuantLaunch {
    const float* audio = nullptr;  // [n_channels][af_stride]
    size_t af_stride = 0;
    uint32_t n_channels = 0;
    uint32_t write_index = 0;  // samples demodulated this slot
    uint32_t af_size = 0;      // (period+5 s)*12000
    uint32_t cover = 0;        // columns [0, cover) of `out` are written (>= write_index: the rest is known to be zero)
    uint32_t out_pitch = 0;    // samples between rows of `out`: af_size, or write_index for the PACKED hand-off layout
                               // ([n_channels][write_index], no zero tail: leaves as one 1-D copy); 0 = af_size
    const unsigned* maxbits = nullptr;
    const float* scale = nullptr;  // [n_channels]
    int16_t* out = nullptr;        // [n_channels][af_size]
    float* factor_out = nullptr;   // [n_channels] final factor (for the log line / stats)
    float* max_out = nullptr;      // [n_channels] maxVal of prepareAudio
};

// Time segmentation of the FAST-tolerance kernels. A slot is cut into SEGMENTS of kFastTile*L - 32 outputs at fixed
// slot-relative positions; L depends on the size of the slot group only (never on how the IQ was pushed), launches of
// the FAST / STFT modes start on segment boundaries, and everything a segment computes is a function of the slot's
// IQ and the slot-relative block index alone -> equal IQ gives equal bytes however it is chunked.
constexpr uint32_t kFastTile = 512;            // SSBD blocks per tile of demod_fast_kernel (128 threads x 4 blocks)
constexpr uint32_t kSegLargeGroupTiles = 3;    // groups of >= kSegLargeGroup channels: 1504 outputs per segment
constexpr uint32_t kSegSmallGroupTiles = 1;    // smaller groups: 480 outputs per segment (more CTAs per launch)
constexpr uint32_t kSegLargeGroup = 64;
uint32_t fast_tiles_per_seg(uint32_t n_group_channels);  // CWSL_TILES_PER_SEG overrides (tests)
inline uint32_t fast_seg_blocks(uint32_t tiles_per_seg) { return tiles_per_seg * kFastTile - 32; }

// Work item of the indirect FAST launch (dynamic-range guard of the STFT mode): <= 32 channels of one segment.
struct GuardItem {
    uint32_t seg;    // segment index inside the launch
    uint32_t first;  // first entry of the segment's selection list
    uint32_t count;  // 1..32
    uint32_t pad;
};
// Channels that a FAST launch takes from device-resident lists instead of the grid (nullptr members = direct launch)
struct FastIndirect {
    const GuardItem* items = nullptr;
    const unsigned* n_items = nullptr;  // device counter written by guard_select_kernel
    const uint32_t* sel = nullptr;      // [n_seg][sel_stride] channel indices, ascending per segment
    uint32_t sel_stride = 0;
    // Phase-table residency: with anchors != nullptr the kernel never reads DemodLaunch::phase. Every CTA replays the
    // reference's exact float recurrence (source/SSBD.hpp:174) from the anchors P_c[128 a] for the <= 32 channels and
    // 512 blocks of the tile it is about to process, into its own slice of `scratch` -- the same bits as the full
    // table, at 1/128 of its memory plus a fixed 39 MB of scratch per stream.
    const float2* anchors = nullptr;    // [anchor][anchor_stride]
    uint32_t anchor_stride = 0;
    const float2* pinc = nullptr;       // [n_channels] the reference's float phase_inc
    float2* scratch = nullptr;          // [grid][kFastGMax][kFastTile], see fast_scratch_bytes()
};
size_t fast_scratch_bytes(int device);  // scratch of one indirect FAST launch on this device

// Extra inputs of the STFT channelizer kernel (cwsl_chan.cu); tables from cwsl_tables.hpp chan_*.
constexpr int kChanTaps = 8;                   // stencil bins per channel of the host-side description (cwsl_stft_channel)
constexpr int kChanKernelWidth = 7;            // Kaiser-Bessel interpolation kernel: support of 7 grid bins
constexpr uint32_t kChanAnchorHops = 128;      // the exact phase recurrence is re-read every 128 hops (slot-relative)
// Interpolation work item: up to 4 channels that are neighbours on the FFT grid read ONE 12-bin window; member j takes
// the 9 bins shift[j] .. shift[j]+8 of it, shift = {0,0,1,2} (cwsl_tables.hpp chan_items). One item per interpolation
// thread, held in registers for the whole launch. Neighbouring items start ~4 bins = two 16-byte bank groups apart, so
// the host deals the items out over the threads such that the eight lanes of every quarter-warp start in eight different
// bank groups (cwsl_tables.hpp chan_item_order): their LDS.128 of the window are then conflict-free.
constexpr int kChanItemMembers = 4, kChanItemTaps = 9, kChanItemBins = 12;
constexpr uint32_t kChanMaxItems = 256;        // per launch
struct alignas(16) ChanItem {                  // 256 bytes, sixteen 16-byte loads
    uint32_t bin_off;                          // byte offset of the window's first bin in a hop buffer: (e mod N) * 8
    uint32_t n_members;
    uint32_t pad0[2];
    uint32_t ch[kChanItemMembers];             // channel index in the slot group, 0xffffffff: empty slot
    float sign[kChanItemMembers];              // +1 USB / -1 LSB
    float pinc[kChanItemMembers][2];           // the reference's float phase_inc (per-hop NCO step between anchors)
    float rot[kChanItemMembers][2];            // e^{-i 240 w_c}
    float w[kChanItemMembers][kChanItemTaps];  // real interpolation weights, exact zeros outside the kernel's support
};
static_assert(sizeof(ChanItem) == 256, "ChanItem layout");
struct ChanLaunch {
    const float* window = nullptr;    // [512]  taps / psihat((j-256)/1024)
    const float2* twiddle = nullptr;  // [32][32] W1024^(j2*q1) * i^q1 at [q1*32 + j2]
    const ChanItem* items = nullptr;  // [n_items]
    uint32_t n_items = 0;             // <= kChanMaxItems
    // P_c[128 a] of the exact float recurrence, [anchor][channel]
    const float2* anchors = nullptr;
    uint32_t anchor_stride = 0;       // channels per anchor row (the whole slot group)
    // Dynamic-range guard statistics per (segment of this launch, channel); seg_blocks == 0: guard off, max|x| goes
    // straight to DemodLaunch::maxbits
    uint32_t seg_blocks = 0;
    unsigned* seg_max = nullptr;      // [n_seg][stat_stride] bit pattern of max|y| over the segment
    unsigned* seg_energy = nullptr;   // [n_seg][stat_stride] sum over octets of min(2^20, sum_8 y^2 * seg_scale[seg])
    const float* seg_scale = nullptr; // [n_seg] 128 / (T^2 * mean|x|^2 of the segment's IQ)
    uint32_t stat_stride = 0;
};
// Guard launch parameters (cwsl_guard.cu)
struct GuardLaunch {
    uint32_t seg_blocks = 0;          // W
    float t2 = 0.0f;                  // T^2: a channel segment keeps the STFT result iff mean y^2 >= T^2 * mean |x|^2
    float* seg_scale = nullptr;       // [n_seg]
    unsigned* seg_max = nullptr;
    unsigned* seg_energy = nullptr;
    uint32_t stat_stride = 0;
    uint32_t* sel = nullptr;          // [n_seg][sel_stride]
    uint32_t sel_stride = 0;
    GuardItem* items = nullptr;       // [n_seg * ceil(C/32)]
    unsigned* n_items = nullptr;
    unsigned long long* counters = nullptr;  // [0] channel-segments decided, [1] of those re-run by the FAST kernel (this slot)
};

// Upload the normalised low-pass taps for one block size into constant memory.
cudaError_t upload_taps(uint32_t block_size, const float* taps /*[32*block_size]*/);
// The compile-time taps the fast kernel uses as immediates, transposed [m*32+n]; nullptr if unsupported.
const float* baked_taps_transposed(uint32_t block_size);

// phase[k] = phase_inc^k by the reference's unfused float recurrence, one thread per table.
cudaError_t launch_phase_tables(const float2* phase_inc /*[n]*/, float2* const* tables /*[n]*/, uint32_t n,
                                uint32_t length, cudaStream_t s);

cudaError_t launch_demod_exact(const DemodLaunch& p, cudaStream_t s);         // tiled, production
cudaError_t launch_demod_exact_gather(const DemodLaunch& p, cudaStream_t s);  // one thread per output, cross-check
// FAST: p.b0 must be a multiple of fast_seg_blocks(tiles_per_seg) (segments sit at fixed slot-relative positions).
// ind.items != nullptr: the channels come from the guard's device-resident work list (persistent grid).
cudaError_t launch_demod_fast(const DemodLaunch& p, uint32_t tiles_per_seg, const FastIndirect& ind, cudaStream_t s);
// STFT channelizer (<= kChanMaxItems work items per launch; DemodLaunch names the whole slot group); p.b0 must be a
// multiple of 32
cudaError_t launch_demod_chan(const DemodLaunch& p, const ChanLaunch& c, cudaStream_t s);
// P_c[spacing * a], a < n_anchor, by the reference's unfused float recurrence (one thread per channel, sequential by
// construction) into [n_anchor][n_channels]: the exact checkpoints of every channel's phase table
cudaError_t launch_phase_anchors(const float2* phase_inc, float2* anchors, uint32_t n_channels, uint32_t n_anchor,
                                 uint32_t spacing, cudaStream_t s);
// Guard: per-segment band power -> seg_scale (before the channelizer), selection + max merge (after it)
cudaError_t launch_guard_band_power(const DemodLaunch& p, const GuardLaunch& g, cudaStream_t s);
cudaError_t launch_guard_select(const DemodLaunch& p, const GuardLaunch& g, cudaStream_t s);
cudaError_t launch_quantise(const QuantLaunch& p, cudaStream_t s);
// p[0..n) = 0; if counters != nullptr: counters[2..3] = counters[0..1] (last finished slot), counters[0..1] = 0
cudaError_t launch_clear_u32(unsigned* p, uint32_t n, unsigned long long* counters, cudaStream_t s);

// FP32 pipe microbenchmark (TFLOP/s).
cudaError_t measure_fp32_peak(float* ffma_tflops, float* ffma2_tflops);

}  // namespace cwsl
