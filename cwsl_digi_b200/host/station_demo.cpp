// station_demo: the reference's main() wiring (source/CWSL_DIGI.cpp:1065-1188) reduced to the
// receive front-end, on synthetic IQ and accelerated time.
//   station_demo <config.ini> <out_dir> [slots=2] [exact|fast]
// Reads the reference's config.ini format, creates one synthetic Receiver per band that has a
// decoder (LO = band centre rounded to 100 kHz), attaches every decoder, then streams IQ as fast
// as the GPU takes it, firing each mode's SyncPredicate every `period` seconds of SIGNAL time, and
// writes the WAV files DecoderPool hands to jt9/wsprd. Needs a CUDA device (no CPU path).
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>

#include "cwsl_host.hpp"

namespace {
// A slot clock in signal time: the receiver's reader thread asks the source for blocks; the source
// fires predicates when the sample count crosses a period boundary (what waitForTime* does on the
// wall clock, source/CWSL_DIGI.cpp:174-451).
class ClockedSource : public SyntheticIqSource {
public:
    ClockedSource(std::uint32_t fs, std::uint32_t iq_len, FrequencyHz lo, std::vector<Carrier> c, std::uint64_t seed,
                  std::uint64_t max_blocks, std::shared_ptr<SyncPredicates> preds, std::set<float> periods)
        : SyntheticIqSource(fs, iq_len, lo, std::move(c), 300.0, seed, max_blocks), fs_(fs), iq_len_(iq_len),
          preds_(std::move(preds)), periods_(std::move(periods)) {}
    bool readBlock(float* dst) override {
        const double t0 = static_cast<double>(n_) / fs_, t1 = static_cast<double>(n_ + iq_len_) / fs_;
        for (float p : periods_)
            if (std::floor(t1 / p) > std::floor(t0 / p) || n_ == 0) preds_->fire(p);
        n_ += iq_len_;
        return SyntheticIqSource::readBlock(dst);
    }

private:
    std::uint32_t fs_, iq_len_;
    std::shared_ptr<SyncPredicates> preds_;
    std::set<float> periods_;
    std::uint64_t n_ = 0;
};
}  // namespace

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <config.ini> <out_dir> [slots] [exact|fast|stft]\n", argv[0]);
        return 2;
    }
    const int slots = argc > 3 ? std::atoi(argv[3]) : 2;
    const std::string modeArg = argc > 4 ? argv[4] : "";   // overrides [gpu] arithmetic= of the config file
    int mode = -1;
    if (!modeArg.empty()) mode = modeArg == "exact" ? CWSL_MODE_EXACT : modeArg == "stft" ? CWSL_MODE_STFT : CWSL_MODE_FAST;
    auto printer = std::make_shared<ScreenPrinter>(LOG_LEVEL::INFO);
    FrontEndConfig cfg;
    try {
        cfg = loadFrontEndConfigFile(argv[1]);
    } catch (const std::exception& e) {
        printer->err(e.what());
        return EXIT_FAILURE;
    }
    if (mode < 0) mode = cfg.kernelMode;
    printer->print("Found " + std::to_string(cfg.decoders.size()) + " decoder entries");
    if (cwsl_device_count() <= 0) {
        printer->err("no CUDA device: the B200 front-end has no CPU path");
        return EXIT_FAILURE;
    }

    std::size_t handled = 0;
    std::mutex mu;
    auto pool = std::make_shared<DecoderPool>("wavefile", argv[2], 4, 300, printer,
                                              [&](const ItemToDecode& it, const std::string& path) {
                                                  std::lock_guard<std::mutex> lk(mu);
                                                  ++handled;
                                                  std::printf("  -> %s %u Hz slot@%llu: %zu samples -> %s\n", it.mode.c_str(),
                                                              it.baseFreq, (unsigned long long)it.epochTime, it.audio.size(),
                                                              path.c_str());
                                              });
    pool->init();

    // one receiver per 192 kHz band segment (findBand(), source/CWSL_Utils.hpp:27-53, picks the CWSL
    // band containing the frequency; here bands are synthetic: LO = freq rounded to 100 kHz)
    // signal time is simulated per receiver, so each receiver gets its own set of slot clocks
    // (the reference has one wall clock for all, source/CWSL_DIGI.cpp:1134-1175)
    std::map<FrequencyHz, std::shared_ptr<SyncPredicates>> preds;
    std::map<FrequencyHz, std::shared_ptr<Receiver>> receivers;
    std::map<FrequencyHz, std::set<float>> periods;
    std::map<FrequencyHz, std::vector<SyntheticIqSource::Carrier>> carriers;
    float longest = 0;
    for (auto& d : cfg.decoders) {
        const FrequencyHz lo = (d.getFreqCalibrated() + 50000) / 100000 * 100000;
        periods[lo].insert(d.getTRPeriod());
        carriers[lo].push_back({d.getFreqCalibrated() + 1500.0, 8000.0});
        longest = std::max(longest, d.getTRPeriod());
    }
    const std::uint32_t fs = 192000, iq_len = 2048;
    const std::uint64_t max_blocks = static_cast<std::uint64_t>(slots * longest * fs / iq_len) + 2;
    int ridx = 0;
    for (auto& kv : periods) {
        preds[kv.first] = std::make_shared<SyncPredicates>();
        auto src = std::make_unique<ClockedSource>(fs, iq_len, kv.first, carriers[kv.first], 20261017 + ridx, max_blocks,
                                                   preds[kv.first], kv.second);
        auto r = std::make_shared<Receiver>("SYNTH" + std::to_string(kv.first / 1000) + "kHz", printer, std::move(src),
                                            ridx % cwsl_device_count(), mode);
        if (!r->init()) return EXIT_FAILURE;
        receivers[kv.first] = r;
        ++ridx;
    }
    std::size_t id = 0;
    for (auto& d : cfg.decoders) {  // setupDecoder(), source/CWSL_DIGI.cpp:103-172
        const FrequencyHz lo = (d.getFreqCalibrated() + 50000) / 100000 * 100000;
        auto inst = std::make_unique<Instance>(receivers[lo], id++, preds[lo]->createPredicate(d.getMode()), d.getFreq(),
                                               d.getFreqCalibrated(), d.getMode(), d.getReporterCallsign(), Wave_SR,
                                               cfg.ftAudioScaleFactor, cfg.wsprAudioScaleFactor, printer, pool,
                                               d.getTRPeriod());
        if (!inst->init()) return EXIT_FAILURE;
        d.setInstance(std::move(inst));
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (auto& kv : receivers) kv.second->start();
    for (auto& kv : receivers) {
        while (kv.second->getStatus() == ReceiverStatus::RUNNING) std::this_thread::sleep_for(std::chrono::milliseconds(5));
        kv.second->terminate();
    }
    pool->drain();
    pool->terminate();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::uint64_t blocks = 0;
    for (auto& kv : receivers) blocks += kv.second->blocksRead();
    std::printf("station_demo: %zu receivers, %zu decoders, %llu IQ blocks (%.1f s of signal per receiver) in %.2f s wall; "
                "%zu audio buffers handed to the decoder pool\n",
                receivers.size(), cfg.decoders.size(), (unsigned long long)blocks, slots * longest, sec, handled);
    return handled > 0 ? EXIT_SUCCESS : EXIT_FAILURE;
}
