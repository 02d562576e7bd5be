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
    const unsigned* maxbits = nullptr;
    const float* scale = nullptr;  // [n_channels]
    int16_t* out = nullptr;        // [n_channels][af_size]
    float* factor_out = nullptr;   // [n_channels] final factor (for the log line / stats)
    float* max_out = nullptr;      // [n_channels] maxVal of prepareAudio
};

// Extra inputs of the STFT channelizer kernel (cwsl_chan.cu); tables from cwsl_tables.hpp chan_*.
constexpr int kChanTaps = 8;                   // stencil bins per channel: Kaiser-Bessel kernel of width 7 on an even-aligned stencil
constexpr int kChanKernelWidth = 7;
constexpr uint32_t kChanMaxChannels = 1024;    // per launch (each interpolation thread keeps <= 4 channels in registers)
struct alignas(16) ChanConst {   // per-channel constants, four 16-byte loads
    int q0;           // first grid bin of the interpolation stencil, even (bins are taken mod 1024)
    float sign;       // +1 USB / -1 LSB
    float rot[2];     // e^{-i 240 w_c}
    float wgt[kChanTaps];  // real interpolation weights of bins q0 .. q0+7
    float pinc[2];    // the reference's float phase_inc (per-hop NCO step between phase-table anchors)
    float pad[2];
};
struct ChanLaunch {
    const float* window = nullptr;    // [512]  taps / psihat((j-256)/1024)
    const float2* twiddle = nullptr;  // [32][32] W1024^(j2*q1) * i^q1 at [q1*32 + j2]
    const ChanConst* consts = nullptr;  // [n_channels]
    int taps = kChanTaps;
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
cudaError_t launch_demod_fast(const DemodLaunch& p, cudaStream_t s);
// STFT channelizer (192 kHz receivers, <= kChanMaxChannels channels per launch)
cudaError_t launch_demod_chan(const DemodLaunch& p, const ChanLaunch& c, cudaStream_t s);
cudaError_t launch_quantise(const QuantLaunch& p, cudaStream_t s);
cudaError_t launch_clear_u32(unsigned* p, uint32_t n, cudaStream_t s);

// FP32 pipe microbenchmark (TFLOP/s).
cudaError_t measure_fp32_peak(float* ffma_tflops, float* ffma2_tflops);

}  // namespace cwsl
