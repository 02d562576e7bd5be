// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin C-ABI wrapper that compiles the REFERENCE's own hot-path headers
// (/root/reference/source/SSBD.hpp + LowPass.hpp, included where they lie, never
// copied) into oracle/_ref/libcwsl_ref*.so, and restates the ~40 arithmetic lines
// of Instance.cpp that cannot be compiled here (windows.h / Qt / boost):
//   demod loop       source/Instance.cpp:259-277
//   prepareAudio     source/Instance.cpp:294-338
//   int16 quantise   source/Instance.cpp:238-241
//   slot reset       source/Instance.cpp:251   (fresh SSBD<float> per slot)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.
//
// Build (oracle/Makefile): g++ -O2 -std=c++17 -ffp-contract=off  (strict = parity oracle)
//                          g++ -O3 -mavx2 -mfma -ffast-math      (speed baseline only)
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>
#include <chrono>
#include <complex>
#include <atomic>

// SSBD keeps filter/tone/phase_inc private; the oracle needs to read them to pin
// the table builder of the product. Test-only access hack, does not alter layout.
#define private public
#include "SSBD.hpp"   // -I/root/reference/source
#undef private

namespace {
constexpr size_t kWaveSR = 12000;   // source/CWSL_DIGI.hpp:51
constexpr size_t kSSBBW  = 6000;    // source/CWSL_DIGI.hpp:52
constexpr bool   kUSB    = true;    // source/CWSL_DIGI.hpp:53
const float kAudioClip = std::pow(2.0f, 15.0f) - 1.0f;  // source/CWSL_DIGI.hpp:55

// source/Instance.cpp:259-277 for a whole slot worth of blocks
size_t demod_slot(uint32_t fs, int32_t demod_freq, const std::complex<float>* iq, size_t n_iq,
                  size_t iq_len, float* af, size_t af_size, bool usb = kUSB) {
    SSBD<float> ssbd(fs, kSSBBW, static_cast<float>(demod_freq), usb);
    const size_t dec_ratio = fs / kWaveSR;               // Instance.cpp:192
    const size_t in_size = ssbd.GetInSize();
    size_t write_index = 0;
    for (size_t blk = 0; blk + iq_len <= n_iq; blk += iq_len) {
        if (write_index + iq_len > af_size - 1) continue;      // Instance.cpp:268-271 ("af buffer full")
        const std::complex<float>* xc = iq + blk;
        float* dest = af + write_index;
        for (size_t n = 0; n < iq_len; n += in_size) ssbd.Iterate(xc + n, dest + n / dec_ratio);
        write_index += iq_len / dec_ratio;
    }
    return write_index;
}

// source/Instance.cpp:294-338
void prepare_audio(float* buf, size_t size, float scale, float* max_out, float* factor_out) {
    float maxVal = std::numeric_limits<float>::lowest();
    for (size_t k = 0; k < size; ++k) if (buf[k] > maxVal) maxVal = buf[k];
    float minVal = (std::numeric_limits<float>::max)();
    for (size_t k = 0; k < size; ++k) if (buf[k] < minVal) minVal = buf[k];
    if (std::fabs(minVal) > maxVal) maxVal = std::fabs(minVal);
    float factor = kAudioClip / (maxVal + 1.0f);
    factor *= scale;
    for (size_t k = 0; k < size; ++k) buf[k] *= factor;
    if (max_out) *max_out = maxVal;
    if (factor_out) *factor_out = factor;
}
}  // namespace

extern "C" {

// Whole reference chain for one decoder and one slot. af (float, af_size) is zeroed
// first like Instance.cpp:213. Returns write_index, or (size_t)-1 if SSBD threw.
size_t cwsl_ref_slot(uint32_t fs, int32_t demod_freq, const float* iq_interleaved, size_t n_iq,
                     size_t iq_len, float scale, size_t af_size, float* af_raw /*nullable*/,
                     int16_t* out_i16, float* max_out, float* factor_out) {
    try {
        std::vector<float> af(af_size, 0.0f);
        const auto* iq = reinterpret_cast<const std::complex<float>*>(iq_interleaved);
        const size_t wi = demod_slot(fs, demod_freq, iq, n_iq, iq_len, af.data(), af_size);
        if (af_raw) std::memcpy(af_raw, af.data(), af_size * sizeof(float));
        prepare_audio(af.data(), af_size, scale, max_out, factor_out);
        for (size_t k = 0; k < af_size; ++k)
            out_i16[k] = static_cast<int16_t>(af[k] + 0.5f);            // Instance.cpp:238-241
        return wi;
    } catch (const std::exception&) {
        return static_cast<size_t>(-1);
    }
}

// Same chain with the sideband selectable (the application always passes USB, source/CWSL_DIGI.hpp:53;
// SSBD itself supports LSB, source/SSBD.hpp:48,110).
size_t cwsl_ref_slot_sb(uint32_t fs, int32_t demod_freq, int is_usb, const float* iq_interleaved, size_t n_iq,
                        size_t iq_len, float scale, size_t af_size, float* af_raw, int16_t* out_i16,
                        float* max_out, float* factor_out) {
    try {
        std::vector<float> af(af_size, 0.0f);
        const auto* iq = reinterpret_cast<const std::complex<float>*>(iq_interleaved);
        const size_t wi = demod_slot(fs, demod_freq, iq, n_iq, iq_len, af.data(), af_size, is_usb != 0);
        if (af_raw) std::memcpy(af_raw, af.data(), af_size * sizeof(float));
        prepare_audio(af.data(), af_size, scale, max_out, factor_out);
        for (size_t k = 0; k < af_size; ++k) out_i16[k] = static_cast<int16_t>(af[k] + 0.5f);
        return wi;
    } catch (const std::exception&) {
        return static_cast<size_t>(-1);
    }
}

// Tables the reference actually uses for (fs, demod_freq): filter[FiltOrder], tone[BlockSize],
// phase_inc. Returns FiltOrder, or 0 if the ctor threw (out-of-band, SSBD.hpp:100-103).
size_t cwsl_ref_tables(uint32_t fs, int32_t demod_freq, int is_usb, float* filter /*>=FiltOrder*/,
                       float* tone_interleaved /*2*BlockSize*/, float* phase_inc2, float* raw_tap_sum) {
    try {
        SSBD<float> s(fs, kSSBBW, static_cast<float>(demod_freq), is_usb != 0);
        for (size_t n = 0; n < s.FiltOrder; ++n) filter[n] = s.filter[n];
        for (size_t n = 0; n < s.BlockSize; ++n) {
            tone_interleaved[2 * n] = s.tone[n].real();
            tone_interleaved[2 * n + 1] = s.tone[n].imag();
        }
        phase_inc2[0] = s.phase_inc.real();
        phase_inc2[1] = s.phase_inc.imag();
        if (raw_tap_sum) {                                   // SSBD.hpp:66-67 on a fresh LowPass
            float* f = BuildLowPass<float>(s.FiltOrder, kSSBBW / (double)fs);
            float sum = 0.0;
            for (size_t n = 0; n < s.FiltOrder; sum += f[n++]);
            *raw_tap_sum = sum;
            delete[] f;
        }
        return s.FiltOrder;
    } catch (const std::exception&) {
        return 0;
    }
}

// The reference's phase after n_blocks ProcessBlock calls (SSBD.hpp:174), by running it.
void cwsl_ref_phase_after(uint32_t fs, int32_t demod_freq, size_t n_blocks, float* phase2) {
    SSBD<float> s(fs, kSSBBW, static_cast<float>(demod_freq), kUSB);
    std::vector<std::complex<float>> zeros(s.GetInSize());
    float out[4];
    for (size_t b = 0; b + 4 <= n_blocks; b += 4) s.Iterate(zeros.data(), out);
    phase2[0] = s.phase.real();
    phase2[1] = s.phase.imag();
}

int cwsl_ref_getters(uint32_t fs, size_t* v /*[6]: InRate OutRate InSize OutSize Bandwidth Delay*/) {
    try {
        SSBD<float> s(fs, kSSBBW, 0.0, kUSB);
        v[0] = s.GetInRate(); v[1] = s.GetOutRate(); v[2] = s.GetInSize();
        v[3] = s.GetOutSize(); v[4] = s.GetBandwidth(); v[5] = s.GetDelay();
        return 0;
    } catch (const std::exception&) { return -1; }
}

// CPU baseline, threaded exactly like the reference: one std::thread per decoder
// (source/Instance.cpp:173), all reading the same IQ. Returns wall seconds start->join.
// out_i16 (n_ch * af_size) may be NULL (outputs discarded after a checksum so the work is not elided).
double cwsl_ref_chain_threads(uint32_t fs, const int32_t* demod_freqs, const float* scales, size_t n_ch,
                              const float* iq_interleaved, size_t n_iq, size_t iq_len, size_t af_size,
                              int16_t* out_i16, size_t max_threads, uint64_t* checksum_out) {
    if (max_threads == 0) max_threads = std::thread::hardware_concurrency();
    if (max_threads == 0) max_threads = 1;
    if (max_threads > n_ch) max_threads = n_ch;
    std::vector<uint64_t> sums(n_ch, 0);
    std::atomic<size_t> next{0};
    const auto t0 = std::chrono::steady_clock::now();
    // max_threads workers, each demodulating whole channels (one channel = one reference
    // Instance thread's work for the slot); with n_ch <= max_threads this is exactly the
    // reference's thread-per-decoder layout, beyond that it avoids oversubscription.
    std::vector<std::thread> th;
    for (size_t w = 0; w < max_threads; ++w) {
        th.emplace_back([&]() {
            std::vector<int16_t> local;
            for (;;) {
                const size_t c = next.fetch_add(1);
                if (c >= n_ch) break;
                int16_t* dst = out_i16 ? out_i16 + c * af_size : (local.resize(af_size), local.data());
                float mx, fac;
                cwsl_ref_slot(fs, demod_freqs[c], iq_interleaved, n_iq, iq_len, scales[c], af_size,
                              nullptr, dst, &mx, &fac);
                uint64_t s = 0;
                for (size_t k = 0; k < af_size; ++k) s = s * 1315423911u + (uint16_t)dst[k];
                sums[c] = s;
            }
        });
    }
    for (auto& t : th) t.join();
    const auto t1 = std::chrono::steady_clock::now();
    uint64_t total = 0;
    for (auto s : sums) total ^= s;
    if (checksum_out) *checksum_out = total;
    return std::chrono::duration<double>(t1 - t0).count();
}

unsigned cwsl_ref_hardware_concurrency(void) { return std::thread::hardware_concurrency(); }

}  // extern "C"
