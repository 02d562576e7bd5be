// Where a Receiver gets its IQ blocks from.
//
// In the reference the only producer is CWSL's Win32 named shared memory (CW Skimmer Server +
// CWSL_Tee): header SM_HDR{SampleRate, BlockInSamples, L0}, one block per WaitForNewData()
// (source/SharedMemory.h:10-21, source/Receiver.hpp:76-97, :209-276). That IPC is Windows-only and
// out of scope; IqSource is the seam a Linux/CWSL producer plugs into. The
// synthetic source below is what the tests, the demo and the benchmarks use.
#pragma once

#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "CWSL_DIGI_Types.hpp"
#include "WsprSynth.hpp"

class IqSource {
public:
    virtual ~IqSource() = default;
    virtual bool open(const std::string& smname) = 0;  // SM.Open(), source/Receiver.hpp:77
    virtual std::uint32_t sampleRate() const = 0;      // SM_HDR.SampleRate
    virtual std::uint32_t blockInSamples() const = 0;  // SM_HDR.BlockInSamples
    virtual FrequencyHz L0() const = 0;                // SM_HDR.L0
    // Blocks until the next block is available (SM.WaitForNewData + SM.Read) and writes
    // blockInSamples() interleaved float32 (I,Q) pairs. false = producer gone (timeout).
    virtual bool readBlock(float* dst) = 0;
};

// Deterministic noise + carriers + M-FSK transmissions with valid codewords (WsprSynth.hpp); unpaced (returns as
// fast as it is read).
class SyntheticIqSource : public IqSource {
public:
    struct Carrier {
        double rf_hz;  // absolute RF frequency
        double amplitude;
    };
    SyntheticIqSource(std::uint32_t fs, std::uint32_t iq_len, FrequencyHz lo, std::vector<Carrier> carriers,
                      double sigma = 300.0, std::uint64_t seed = 20261017, std::uint64_t max_blocks = ~0ull)
        : fs_(fs), iq_len_(iq_len), lo_(lo), carriers_(std::move(carriers)), sigma_(sigma), state_(seed),
          max_blocks_(max_blocks) {}
    void addBurst(FskBurst b) { bursts_.push_back(std::move(b)); }  // before the first readBlock()
    bool open(const std::string&) override { return true; }
    std::uint32_t sampleRate() const override { return fs_; }
    std::uint32_t blockInSamples() const override { return iq_len_; }
    FrequencyHz L0() const override { return lo_; }
    bool readBlock(float* dst) override {
        if (blocks_ >= max_blocks_) return false;
        const double two_pi = 2.0 * 3.14159265358979323846;
        for (std::uint32_t i = 0; i < iq_len_; ++i) {
            double re, im;
            gauss(re, im);
            re *= sigma_;
            im *= sigma_;
            for (const Carrier& c : carriers_) {
                const double f = c.rf_hz - static_cast<double>(lo_);
                const double cyc = std::fmod(f * static_cast<double>(n_ % fs_) / fs_, 1.0);
                re += c.amplitude * std::cos(two_pi * cyc);
                im += c.amplitude * std::sin(two_pi * cyc);
            }
            for (FskBurst& b : bursts_) {  // continuous phase: the frequency is integrated sample by sample
                double t = static_cast<double>(n_) / fs_;
                if (b.period_s > 0) t = std::fmod(t, b.period_s);
                const double k = std::floor((t - b.t0_s) / b.symbol_s);
                if (t < b.t0_s || k >= static_cast<double>(b.symbols.size())) continue;
                const double f = b.rf_hz + b.tone_hz * b.symbols[static_cast<std::size_t>(k)] - static_cast<double>(lo_);
                b.phase += f / fs_;
                b.phase -= std::floor(b.phase);
                re += b.amplitude * std::cos(two_pi * b.phase);
                im += b.amplitude * std::sin(two_pi * b.phase);
            }
            dst[2 * i] = static_cast<float>(re);
            dst[2 * i + 1] = static_cast<float>(im);
            ++n_;
        }
        ++blocks_;
        return true;
    }

private:
    std::uint64_t next() {  // SplitMix64
        std::uint64_t z = (state_ += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    void gauss(double& a, double& b) {  // Box-Muller
        const double u1 = (static_cast<double>(next() >> 11) + 1.0) / 9007199254740993.0;
        const double u2 = static_cast<double>(next() >> 11) / 9007199254740992.0;
        const double r = std::sqrt(-2.0 * std::log(u1));
        a = r * std::cos(2.0 * 3.14159265358979323846 * u2);
        b = r * std::sin(2.0 * 3.14159265358979323846 * u2);
    }
    std::uint32_t fs_, iq_len_;
    FrequencyHz lo_;
    std::vector<Carrier> carriers_;
    std::vector<FskBurst> bursts_;
    double sigma_;
    std::uint64_t state_;
    std::uint64_t max_blocks_;
    std::uint64_t n_ = 0, blocks_ = 0;
};
