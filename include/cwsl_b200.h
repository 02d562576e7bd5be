/*
 * cwsl_b200.h -- C ABI of the B200-native CWSL_DIGI receive front-end.
 *
 * The reference (alexranaldi/CWSL_DIGI v0.88) has no FFI/plugin seam for this path: the seam is
 * a set of C++ classes inside one executable. This header is the boundary a maintainer binds
 * underneath those classes; every entry point names the reference interface it replaces
 * (paths relative to the reference checkout). INTEGRATION.md shows the reference-side glue.
 *
 * Conventions: plain pointers and sizes only; opaque handles; int status (0 = CWSL_OK,
 * negative = error, text via cwsl_last_error()); the caller owns every host buffer, the library
 * owns all device memory; calls on one handle are serialised by the caller, different handles
 * may be driven from different threads / GPUs concurrently. There is no CPU fallback: if no
 * CUDA device is usable every compute entry point fails with CWSL_ERR_CUDA.
 */
#ifndef CWSL_B200_H
#define CWSL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CWSL_B200_ABI_VERSION 3

#define CWSL_OK 0
#define CWSL_ERR_INVALID (-1)  /* bad argument; where the reference throws std::invalid_argument
                                  (source/SSBD.hpp:54-59, :100-103) or rejects a config value */
#define CWSL_ERR_CUDA (-2)     /* CUDA runtime/driver failure, or no device */
#define CWSL_ERR_NOMEM (-3)
#define CWSL_ERR_STATE (-4)    /* call-order error (e.g. push before any channel exists) */
#define CWSL_ERR_OVERRUN (-5)  /* IQ ring overrun: data needed by an open slot was overwritten
                                  (the reference logs "ring buffer full", source/Receiver.hpp:222-229) */

/* Arithmetic mode of the demodulator kernels.
 * EXACT: every float operation of source/SSBD.hpp:160-183 is issued unfused, in the reference's
 *        order -> float audio and int16 output are bit-identical to the reference chain built
 *        with strict IEEE flags (oracle/_ref).
 * FAST:  same tables and the same float phase recurrence, FMA-contracted mix and FIR
 *        (packed fma.rn.f32x2) -> <= 1 int16 LSB, residual >= 90 dB below signal.
 * STFT:  for receivers with many channels: per output sample (hop of one SSBD block) ONE FFT of the windowed last
 *        FiltOrder IQ samples on a 2*FiltOrder grid (1024 / 512 / 256 bins at 192 / 96 / 48 kHz) is shared by every
 *        channel of the slot group; each channel reads its NCO frequency off that grid (8-bin Kaiser-Bessel
 *        interpolation) and applies the reference's own float phase recurrence. Used for groups of >= 64 channels
 *        (CWSL_STFT_MIN_CHANNELS; the measured break-even with the FAST kernel); smaller groups run the FAST kernel.
 *        Contract: the bars of FAST (<= 1 int16 LSB, residual >= 90 dB below the channel's signal wherever FAST
 *        itself reaches them). The FFT's own error is NOT relative to the channel: it is the rounding noise of a
 *        float32 FFT, 0.6e-7 ... 5.8e-7 (-144 ... -125 dB) of the rms of the whole band. The bars are kept by a
 *        dynamic-range guard: per segment of 1504 output samples, a channel whose mean power lies more than 32 dB
 *        (cwsl_rx_set_stft_guard) under the band's is recomputed by the FAST kernel on the device before the slot is
 *        normalised, so every sample handed on is either an FFT sample >= 92 dB above the FFT floor or a FAST sample.
 * FAST and STFT results are functions of the slot's IQ alone: segment and anchor positions are fixed in
 * slot-relative coordinates, so equal IQ gives equal bytes however it was pushed (cwsl_rx_push_iq chunking,
 * cwsl_rx_process calls). For that the two modes demodulate whole segments only until the slot edge. */
#define CWSL_MODE_EXACT 0
#define CWSL_MODE_FAST 1
#define CWSL_MODE_STFT 2

typedef struct cwsl_rx cwsl_rx_t;

/* ---- library / device ------------------------------------------------------------------ */
int cwsl_abi_version(void);
const char* cwsl_last_error(void); /* thread-local text of the last failure on this thread */
int cwsl_device_count(void);       /* number of visible CUDA devices, 0 if none */

/* ---- SSBD parameter algebra (host only, no GPU needed) --------------------------------- */
/* Replaces SSBD<>::GetInRate/GetOutRate/GetInSize/GetOutSize/GetBandwidth/GetDelay
 * (source/SSBD.hpp:140-154) plus the derived FiltOrder/BlockSize/NumWS (:62,:71,:75).
 * out[9] = {InRate, OutRate, InSize, OutSize, Bandwidth, Delay, FiltOrder, BlockSize, NumWS}.
 * CWSL_ERR_INVALID where the SSBD ctor throws (source/SSBD.hpp:54-59). */
int cwsl_ssbd_params(uint32_t sample_rate, uint32_t out[9]);

/* The tables SSBD<float>(Fs, 6000, float(demod_freq_hz), is_usb) builds in its ctor and Tune()
 * (source/SSBD.hpp:62-68, :110-114; source/LowPass.hpp:16-35): normalised filter[FiltOrder],
 * tone[2*BlockSize] (re,im interleaved) and phase_inc[2]. Host only. CWSL_ERR_INVALID for
 * out-of-band tunings (source/SSBD.hpp:100-103). */
int cwsl_build_tables(uint32_t sample_rate, int32_t demod_freq_hz, int is_usb, float* filter,
                      float* tone, float* phase_inc);

/* Host-side constants of CWSL_MODE_STFT, for tests and diagnostics (no device needed). L = FiltOrder, N = 2 L.
 * window[L]        low-pass taps divided by the transform of the interpolation kernel;
 * twiddle[2*N]     (re,im) of W_N^(j2*q1) * i^q1 at [q1*32 + j2], q1 < N/32: the inter-pass factors of the
 *                  (N/32) x 32 FFT;
 * per channel: q0 = first (even) grid bin of the 8-bin stencil, wgt[8] its real interpolation weights,
 * rot[2] = e^{-i (L/2 - BlockSize) w}, w = the channel's NCO step per input sample. Any output pointer may be NULL.
 * audio[b] = Weaver select of  phase[b] * rot * sum_i wgt[i] * X_b[(q0+i) mod N],  X_b[q] = i^q * N-point FFT of
 * (window * the last L IQ samples up to and including SSBD block b). */
int cwsl_stft_tables(uint32_t sample_rate, float* window, float* twiddle);
int cwsl_stft_channel(uint32_t sample_rate, int32_t demod_freq_hz, int is_usb, int32_t* q0, float* wgt,
                      float* rot);
/* How the channelizer kernel groups a channel set (no device needed): channels that are neighbours on the FFT grid are
 * packed, in ascending grid position, into work items of <= 4 members that read one 12-bin window starting at the
 * (even) bin first_bin[i]; member j uses the 9 bins first_bin[i] + shift[j] ... + 8, shift = {0,0,1,2}, with the real
 * weights weights[i][j][0..8] (exact zeros outside the interpolation kernel's support, so every channel meets the
 * same non-zero weights on the same bins as in cwsl_stft_channel). channels[i][j] = index into demod_freq_hz, or -1
 * for an empty slot. Arrays are sized for n items (the worst case: no two channels share a window); *n_items returns
 * how many were built. */
int cwsl_stft_items(uint32_t sample_rate, const int32_t* demod_freq_hz, const int* is_usb, uint32_t n,
                    int32_t* first_bin, int32_t* channels, float* weights, uint32_t* n_items);

/* (period + 5 s) * 12000: length of one decoder's audio buffer, source/Instance.cpp:149. */
size_t cwsl_af_size(double period_s);

/* Number of IQ blocks the demod loop accepts before its "af buffer full" guard drops the rest
 * (source/Instance.cpp:268-271). */
size_t cwsl_accepted_blocks(size_t n_iq_blocks, uint32_t iq_len, uint32_t sample_rate, size_t af_size);

/* ---- receiver: one per CWSL shared-memory band (source/Receiver.hpp:52-302) -------------- */
/* sample_rate / iq_len are SM_HDR.SampleRate / BlockInSamples (source/Receiver.hpp:87-88).
 * iq_len must be a multiple of SSBD::GetInSize() (the reference silently over-reads otherwise,
 * source/Instance.cpp:273). ring_seconds sizes the device-resident IQ ring (the reference keeps
 * ~3 s of blocks on the host, source/Receiver.hpp:132); 0 = size it for the longest slot of the
 * groups added before the first push. It is a minimum: the ring always holds at least two segments
 * (about 0.26 s). Returns NULL on failure (see cwsl_last_error). */
cwsl_rx_t* cwsl_rx_create(int device, uint32_t sample_rate, uint32_t iq_len, double ring_seconds);
void cwsl_rx_destroy(cwsl_rx_t* rx);

/* CWSL_MODE_EXACT, CWSL_MODE_FAST (default) or CWSL_MODE_STFT. Takes effect at each slot group's next slot edge (the
 * mode of a slot is latched at its first demodulation), so a slot is never a mixture of two modes. */
int cwsl_rx_set_mode(cwsl_rx_t* rx, int mode);

/* Threshold of the STFT mode's dynamic-range guard, in dB below the mean power of the band: channel segments whose
 * mean power is lower are recomputed in the direct form (see CWSL_MODE_STFT). Default 32 (CWSL_STFT_GUARD_DB);
 * 0 switches the guard off (diagnostics: the raw FFT channelizer). */
int cwsl_rx_set_stft_guard(cwsl_rx_t* rx, double db_below_band_power);

/* A slot group = the decoders that share one SyncPredicate, i.e. one mode/period
 * (source/CWSL_DIGI_Types.hpp:65-145, source/CWSL_DIGI.cpp:134). period_s as getRXPeriod()
 * (source/CWSL_DIGI.hpp:64-113). Returns the group id (>= 0) or a negative error. */
int cwsl_rx_add_group(cwsl_rx_t* rx, double period_s);

/* One decoder channel = one Instance + its SSBD<float>(Fs, SSB_BW, float(demod_freq_hz), USB)
 * (source/Instance.cpp:183-187). demod_freq_hz = int32(calibratedSSBFreq - LO)
 * (source/Instance.cpp:183); scale = audioScaleFactor_ft or _wspr, 0 < scale <= 1
 * (source/Instance.cpp:320-329, source/CWSL_DIGI.cpp:952-978). Builds the reference's tables on
 * the host and the float phase sequence phase_inc^k (source/SSBD.hpp:174) on the device.
 * Returns the channel index inside the group (>= 0) or a negative error; CWSL_ERR_INVALID for
 * tunings SSBD::Tune rejects (source/SSBD.hpp:100-103).
 * On a receiver that is already running (a decoder restarted or moved to this band by the main loop,
 * source/CWSL_DIGI.cpp:1217-1226 -> setupDecoder) the channel joins at the group's next slot edge -- immediately if
 * no IQ of an open slot is pending; the returned index is the one it will have then (channels keep their order;
 * a removal pending for the same edge moves it down). A cwsl_rx_end_slot that has
 * such a change pending must take its result to the host (out_i16 != NULL): the device buffers are replaced. */
int cwsl_rx_add_channel(cwsl_rx_t* rx, int group, int32_t demod_freq_hz, int is_usb, float scale);

/* Remove a decoder channel (Instance::terminate, source/Instance.cpp:100-119). Before the first push: at once.
 * On a running receiver: at the group's next slot edge (immediately between slots); channels behind it move down
 * by one index then. A group cannot lose its last channel (destroy the receiver instead). */
int cwsl_rx_remove_channel(cwsl_rx_t* rx, int group, int channel);

int cwsl_rx_num_groups(const cwsl_rx_t* rx);
int cwsl_rx_num_channels(const cwsl_rx_t* rx, int group);
size_t cwsl_rx_group_af_size(const cwsl_rx_t* rx, int group);

/* Append n_blocks IQ blocks (n_blocks * iq_len complex samples, interleaved float32 I,Q as in
 * source/Receiver.hpp:140) from HOST memory to the receiver's device ring. Replaces
 * Receiver::readIQ's copy into the ring (source/Receiver.hpp:242-249) AND the N per-decoder
 * pops of the same block (source/Instance.cpp:260-265): the block crosses PCIe once.
 * Asynchronous when `iq` is pinned memory. Slots that would lose un-demodulated samples to the
 * ring wrap are demodulated first. */
int cwsl_rx_push_iq(cwsl_rx_t* rx, const float* iq, size_t n_blocks);

/* Fences for callers that stage IQ in a ring of pinned buffers (source/Receiver.hpp:209-276 keeps such a ring on the
 * host): cwsl_rx_push_fence marks the receiver's stream behind everything pushed so far and returns a token;
 * cwsl_rx_wait_fence blocks until the work in front of that mark -- in particular the host-to-device copies that
 * read the staging buffers -- has completed, without waiting for kernels or copies queued later. At most 64 fences
 * are tracked; waiting for an older one waits for a younger one instead (never too short). */
int cwsl_rx_push_fence(cwsl_rx_t* rx, uint64_t* token);
int cwsl_rx_wait_fence(cwsl_rx_t* rx, uint64_t token);

/* Same, source already in device memory on rx's device (device-to-device copy into the ring).
 * Stream semantics: the copy is queued on the receiver's stream (cwsl_rx_stream), which does not wait for any
 * other stream; whatever produced d_iq must have completed, or be ordered before that stream by the caller. */
int cwsl_rx_push_iq_device(cwsl_rx_t* rx, const float* d_iq, size_t n_blocks);

/* Zero-copy variant for benchmarks: treat the device buffer d_iq (n_blocks * iq_len samples,
 * must stay valid until the slot ends) as the complete IQ of the NEXT slot of every group.
 * Replaces any ring contents; nothing is copied. Same stream semantics as cwsl_rx_push_iq_device. */
int cwsl_rx_bind_device_iq(cwsl_rx_t* rx, const float* d_iq, size_t n_blocks);

/* Demodulate what has been pushed so far for `group` (or all groups if group < 0) without ending
 * the slot: the streaming counterpart of the per-block Iterate loop, source/Instance.cpp:273-276.
 * EXACT mode: everything; FAST / STFT: all complete segments (1504 output samples for groups of >= 64 channels,
 * 480 for smaller ones), the rest waits in the ring for the next call or the slot edge. Asynchronous. */
int cwsl_rx_process(cwsl_rx_t* rx, int group);

/* Slot edge for `group` = the SyncPredicate firing (source/Instance.cpp:203-253): finishes the
 * demodulation of the samples pushed since the previous edge, then for EVERY channel of the
 * group runs prepareAudio (max|x| over the whole buffer, factor = 32767/(max+1)*scale,
 * source/Instance.cpp:294-338) and the int16 conversion (int16)(x+0.5f)
 * (source/Instance.cpp:238-241), and resets the group's demodulators (fresh SSBD per slot,
 * source/Instance.cpp:251). out_i16 is HOST memory [n_channels][af_size] (may be NULL to leave
 * the result on the device); *write_index receives the number of audio samples demodulated
 * into each buffer (the rest of af_size is the zero tail). The host copy is asynchronous when
 * out_i16 is pinned: call cwsl_rx_synchronize() before reading it. */
int cwsl_rx_end_slot(cwsl_rx_t* rx, int group, int16_t* out_i16, size_t* write_index);

/* Same slot edge, PACKED hand-off: out_i16 receives [n_channels][*write_index] -- every channel's demodulated samples
 * back to back, without the zero tail (the consumer knows write_index and pads when it builds the decoder's buffer of
 * (period + 5 s) * 12000 samples, source/Instance.cpp:149, :238-241). out_i16 must hold n_channels * af_size samples
 * (the worst case). The result is one contiguous range in device and host memory, so it crosses PCIe as a single 1-D
 * copy: 55.6 GB/s on B200 / Gen5 x16 where the strided [n_channels][af_size] copy reaches 51.9. After a packed slot
 * cwsl_rx_device_audio() points at the packed layout too; cwsl_rx_copy_device_audio() still delivers af_size samples. */
int cwsl_rx_end_slot_packed(cwsl_rx_t* rx, int group, int16_t* out_i16, size_t* write_index);

/* Pinned, zero-initialised host memory for slot hand-off buffers. When the out_i16 of
 * cwsl_rx_end_slot lies inside such a region the library tracks which columns it has ever written
 * there and copies only those that can be non-zero (the zero tail past write_index is already zero
 * on the host), which removes 25 % of the PCIe traffic of an FT8 slot. The caller must treat the
 * buffer as read-only. Any other host pointer always receives the full [n_channels][af_size] copy.
 * NULL on failure. */
void* cwsl_host_alloc(size_t bytes);
void cwsl_host_free(void* p);

/* Device pointer of the last finished slot's int16 audio, [n_channels][af_size]; valid until
 * the next cwsl_rx_end_slot on the same group. */
const int16_t* cwsl_rx_device_audio(const cwsl_rx_t* rx, int group);

/* Copy one channel's int16 audio of the last finished slot (af_size samples) into the caller's
 * DEVICE buffer, asynchronously on the receiver's stream (e.g. as the send buffer of an NCCL
 * gather of per-slot audio to the station's rank 0). */
int cwsl_rx_copy_device_audio(cwsl_rx_t* rx, int group, int channel, int16_t* d_dst);

/* Float audio before normalisation of the last finished slot (what prepareAudio reads,
 * source/Instance.cpp:295), copied to HOST out[af_size] (zero tail included); synchronous.
 * Parity/debug aid. */
int cwsl_rx_read_float_audio(cwsl_rx_t* rx, int group, int channel, float* out);

/* maxVal and final factor prepareAudio computed for the last finished slot
 * (the values the reference logs, source/Instance.cpp:314,336); synchronous. */
int cwsl_rx_channel_stats(cwsl_rx_t* rx, int group, int channel, float* max_val, float* factor);

/* Block until all asynchronous work queued on this receiver has finished (kernels and copies). */
int cwsl_rx_synchronize(cwsl_rx_t* rx);

/* Block only until the host copy of the LAST cwsl_rx_end_slot(rx, ..., out_i16, ...) has landed. The copy runs
 * on a private copy stream behind the slot's kernels, so it overlaps whatever is queued next on the receiver's
 * stream; use this (not cwsl_rx_synchronize) when several receivers share one stream and only this receiver's
 * hand-off buffer is needed. */
int cwsl_rx_wait_output(cwsl_rx_t* rx);

/* Device-side join, no host blocking: whatever is queued on the receiver's stream after this call starts only when
 * the post work of the last cwsl_rx_end_slot (normalise/quantise, max reset, host copy -- they run on a private
 * post stream) has finished. For callers that bracket several receivers with events on one shared stream. */
int cwsl_rx_join_output(cwsl_rx_t* rx);

/* The CUDA stream (cudaStream_t) all work of this receiver is queued on, for event timing. */
void* cwsl_rx_stream(cwsl_rx_t* rx);

/* Queue all further work of this receiver on the caller's stream (cudaStream_t; NULL = the legacy
 * default stream) instead of the private one, e.g. to order it with the caller's own copies and
 * events. The receiver never destroys a caller-supplied stream. */
int cwsl_rx_set_stream(cwsl_rx_t* rx, void* cuda_stream);

/* Record CUDA events around every kernel launch of this receiver (off by default). */
int cwsl_rx_enable_timing(cwsl_rx_t* rx, int on);

/* Device time (ms, CUDA events on the receiver's stream) of the demodulator kernel launches and
 * of the normalise+quantise launches queued since the last call; also their launch counts.
 * Synchronises the stream. Any pointer may be NULL. */
int cwsl_rx_kernel_times(cwsl_rx_t* rx, float* demod_ms, float* quant_ms, int* demod_launches,
                         int* quant_launches);

/* ms[5] = {demodulation total, normalise+quantise total, main demodulator kernel(s) only, STFT guard prologue
 * (band power), STFT guard epilogue (selection + direct-form redo)}, launches[2] = {demodulation passes,
 * quantise passes} since the last call; otherwise like cwsl_rx_kernel_times. */
int cwsl_rx_kernel_times_ex(cwsl_rx_t* rx, float ms[5], int launches[2]);

/* STFT dynamic-range guard, last finished slot of `group`: channel segments decided, and how many of those were
 * recomputed by the FAST kernel. Synchronous. */
int cwsl_rx_guard_stats(cwsl_rx_t* rx, int group, uint64_t* decided, uint64_t* redone);

/* ---- measurement helpers ---------------------------------------------------------------- */
/* FP32 FMA-pipe peak of `device`, measured with a register-resident FFMA / packed FFMA2
 * loop (TFLOP/s, 2 flop per lane-FMA). Used as the roofline denominator of this path, which
 * is FP32-pipe-bound. */
int cwsl_measure_fp32_peak(int device, float* ffma_tflops, float* ffma2_tflops);

#ifdef __cplusplus
}
#endif
#endif /* CWSL_B200_H */
