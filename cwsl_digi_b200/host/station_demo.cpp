// station_demo: the reference's main() wiring (source/CWSL_DIGI.cpp:1065-1188) reduced to the
// receive front-end, on synthetic IQ and accelerated time.
//   station_demo <config.ini> <out_dir> [slots=2] [exact|fast|stft] [wavefile|shmem] [dumpiq] [shmsource]
// shmsource: the IQ does not come straight from the synthetic source but through the CWSL shared-memory ring
// (row f3): a producer thread writes blocks into a POSIX segment with the reference's layout, the Receiver's reader
// thread takes them out with CwslShmSource into its pinned staging ring and pushes them to the GPU.
// shmem: FT8/FT4/JT65/Q65 items go through jt9's shared-memory block (DecoderPool, row f2) to an in-process stand-in
// for jt9 that attaches by key, saves d2[] + the parameters it was given to <out_dir>/<key>.d2 / .txt and plays
// jt9's side of the ipc[] handshake. dumpiq: every receiver's IQ stream is also written to <out_dir>/iq_<LO>.f32
// together with the IQ-block index of each slot edge (<out_dir>/edges_<LO>.txt), so a test can re-derive every
// hand-off artefact from the same samples with the oracle.
// Reads the reference's config.ini format, creates one synthetic Receiver per band that has a
// decoder (LO = band centre rounded to 100 kHz), attaches every decoder, then streams IQ as fast
// as the GPU takes it, firing each mode's SyncPredicate every `period` seconds of SIGNAL time, and
// writes the WAV files DecoderPool hands to jt9/wsprd. Needs a CUDA device (no CPU path).
#include <cstdio>
#include <cstdlib>
#include <future>
#include <map>
#include <set>

#include "cwsl_host.hpp"

namespace {
// The transmissions every WSPR decoder of the demo hears (audio frequency of the signal's centre, SNR in 2500 Hz,
// start within the slot); tests/test_host_gpu.py expects exactly these messages back from the WAV hand-off.
struct WsprDemoTx {
    const char* call;
    const char* grid;
    int dbm;
    double audio_hz, snr_db, t0_s;
};
const WsprDemoTx kWsprDemoTransmissions[] = {
    {"K1ABC", "FN42", 37, 1500.0, -12.0, 1.0},
    {"W1AW", "FN31", 30, 1440.0, -20.0, 1.3},
    {"G4JNT", "IO90", 23, 1570.0, -24.0, 0.8},
};

// A slot clock in signal time: the receiver's reader thread asks the source for blocks; the source
// fires predicates when the sample count crosses a period boundary (what waitForTime* does on the
// wall clock, source/CWSL_DIGI.cpp:174-451).
class ClockedSource : public IqSource {
public:
    ClockedSource(std::unique_ptr<IqSource> inner, std::shared_ptr<SyncPredicates> preds, std::set<float> periods,
                  const std::string& dumpDir = "")
        : inner_(std::move(inner)), preds_(std::move(preds)), periods_(std::move(periods)), dumpDir_(dumpDir) {}
    ~ClockedSource() override {
        if (iq_) std::fclose(iq_);
        if (edges_) std::fclose(edges_);
    }
    bool open(const std::string& smname) override {
        if (!inner_->open(smname)) return false;
        fs_ = inner_->sampleRate();
        iq_len_ = inner_->blockInSamples();
        if (!dumpDir_.empty()) {
            iq_ = std::fopen((dumpDir_ + "/iq_" + std::to_string(inner_->L0()) + ".f32").c_str(), "wb");
            edges_ = std::fopen((dumpDir_ + "/edges_" + std::to_string(inner_->L0()) + ".txt").c_str(), "w");
        }
        return true;
    }
    std::uint32_t sampleRate() const override { return inner_->sampleRate(); }
    std::uint32_t blockInSamples() const override { return inner_->blockInSamples(); }
    FrequencyHz L0() const override { return inner_->L0(); }
    bool readBlock(float* dst) override {
        const double t0 = static_cast<double>(n_) / fs_, t1 = static_cast<double>(n_ + iq_len_) / fs_;
        for (float p : periods_)
            if (std::floor(t1 / p) > std::floor(t0 / p) || n_ == 0) {
                preds_->fire(p);
                // the Receiver sees the flag at the top of its next loop, i.e. after it has pushed THIS block:
                // the slot that ends holds the IQ blocks up to and including block n_/iq_len_
                if (edges_) std::fprintf(edges_, "%g %llu\n", (double)p, (unsigned long long)(n_ / iq_len_));
            }
        n_ += iq_len_;
        const bool ok = inner_->readBlock(dst);
        if (ok && iq_) std::fwrite(dst, sizeof(float), 2 * iq_len_, iq_);
        return ok;
    }

private:
    std::unique_ptr<IqSource> inner_;
    std::uint32_t fs_ = 0, iq_len_ = 0;
    std::shared_ptr<SyncPredicates> preds_;
    std::set<float> periods_;
    std::string dumpDir_;
    std::uint64_t n_ = 0;
    FILE* iq_ = nullptr;
    FILE* edges_ = nullptr;
};

// CWSL's producer side (CWSL_Tee writing the band's shared-memory ring, source/SharedMemory.cpp:158-202), here a
// thread that copies a synthetic source into a POSIX segment with the reference's byte layout. It never runs more
// than half a ring ahead of the reader (the real producer is paced by the radio's clock instead).
class ShmProducer {
public:
    ShmProducer(const std::string& name, std::unique_ptr<SyntheticIqSource> src, std::uint32_t ringBlocks)
        : name_(name), src_(std::move(src)), ringBlocks_(ringBlocks) {}
    ~ShmProducer() { join(); }
    bool start() {
        const std::uint32_t bytes = src_->blockInSamples() * 8u;
        SM_HDR hdr{(int)src_->sampleRate(), (int)src_->blockInSamples(), (int)src_->L0()};
        if (!sm_.Create(name_, bytes * ringBlocks_, hdr)) return false;
        th_ = std::thread([this, bytes] {
            std::vector<float> blk(bytes / 4);
            while (src_->readBlock(blk.data())) {
                while (written_ - consumed.load() >= ringBlocks_ / 2) std::this_thread::sleep_for(std::chrono::microseconds(50));
                sm_.Write(reinterpret_cast<const std::uint8_t*>(blk.data()), bytes);
                ++written_;
            }
        });
        return true;
    }
    void join() {
        if (th_.joinable()) th_.join();
    }
    std::atomic<std::uint64_t> consumed{0};  // advanced by the reader side of the demo

private:
    std::string name_;
    std::unique_ptr<SyntheticIqSource> src_;
    std::uint32_t ringBlocks_;
    CSharedMemory sm_;
    std::thread th_;
    std::uint64_t written_ = 0;
};

// CwslShmSource + flow control feedback for ShmProducer
class CountingShmSource : public CwslShmSource {
public:
    explicit CountingShmSource(ShmProducer* p) : CwslShmSource(300), prod_(p) {}
    bool readBlock(float* dst) override {
        const bool ok = CwslShmSource::readBlock(dst);
        if (ok) ++prod_->consumed;
        return ok;
    }

private:
    ShmProducer* prod_;
};

// In-process stand-in for "jt9 -s <key>": attaches to the segment by key, takes what a decoder would take (d2[] and
// the parameter block), plays jt9's side of the handshake (clear ipc[1] when done with the data, leave when
// ipc[2] == 1) on its own thread, like the separate process it replaces.
class FakeJt9 {
public:
    explicit FakeJt9(std::string outDir) : dir(std::move(outDir)) {}
    ~FakeJt9() { joinAll(); }
    void joinAll() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& t : threads)
            if (t.joinable()) t.join();
    }
    bool run(const std::string& key, const ItemToDecode& item) {
        std::promise<bool> printed;
        auto fut = printed.get_future();
        std::lock_guard<std::mutex> lk(mu);
        threads.emplace_back([this, key, item, p = std::move(printed)]() mutable {
            Jt9ShmSegment seg;
            if (!seg.attach(key)) {
                p.set_value(false);
                return;
            }
            dec_data_t* d = seg.data();
            const bool sane = d->ipc[1] == 1 && d->ipc[2] == -1 && d->ipc[0] == d->params.nzhsym && d->params.newdat;
            const std::string base = dir + "/" + key;
            if (FILE* f = std::fopen((base + ".d2").c_str(), "wb")) {
                std::fwrite(d->d2, sizeof(short), item.audio.size(), f);
                std::fclose(f);
            }
            if (FILE* f = std::fopen((base + ".txt").c_str(), "w")) {
                std::fprintf(f, "mode=%s instance=%d freq=%u nmode=%d ntrperiod=%d nzhsym=%d ndepth=%d nfb=%d sane=%d\n",
                             item.mode.c_str(), item.instanceId, item.baseFreq, d->params.nmode, d->params.ntrperiod,
                             d->params.nzhsym, d->params.ndepth, d->params.nfb, (int)sane);
                std::fclose(f);
            }
            p.set_value(sane);           // "<DecodeFinished>" reached the pipe
            volatile int* ipc = d->ipc;
            ipc[1] = 0;                  // done with the data
            for (int i = 0; i < 20000 && ipc[2] != 1; ++i) std::this_thread::sleep_for(std::chrono::milliseconds(1));
            terminated += ipc[2] == 1 && ipc[1] == 999;
        });
        return fut.get();
    }
    std::atomic<int> terminated{0};

private:
    std::string dir;
    std::mutex mu;
    std::vector<std::thread> threads;
};
}  // namespace

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <config.ini> <out_dir> [slots] [exact|fast|stft] [wavefile|shmem] [dumpiq]\n", argv[0]);
        return 2;
    }
    const int slots = argc > 3 ? std::atoi(argv[3]) : 2;
    const std::string modeArg = argc > 4 ? argv[4] : "";   // overrides [gpu] arithmetic= of the config file
    int mode = -1;
    if (!modeArg.empty()) mode = modeArg == "exact" ? CWSL_MODE_EXACT : modeArg == "stft" ? CWSL_MODE_STFT : CWSL_MODE_FAST;
    std::string transfer = "wavefile";   // wsjtx.transfermethod, source/CWSL_DIGI.cpp:1023
    std::string dumpDir;
    bool shmSource = false;
    for (int i = 5; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "wavefile" || a == "shmem") transfer = a;
        else if (a == "dumpiq") dumpDir = argv[2];
        else if (a == "shmsource") shmSource = true;
    }
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
    FakeJt9 jt9(argv[2]);
    auto pool = std::make_shared<DecoderPool>(
        transfer, argv[2], 4, 300, printer,
        [&](const ItemToDecode& it, const std::string& path) {
            std::lock_guard<std::mutex> lk(mu);
            ++handled;
            std::printf("  -> %s %u Hz slot@%llu: %zu samples -> %s\n", it.mode.c_str(), it.baseFreq,
                        (unsigned long long)it.epochTime, it.audio.size(), path.empty() ? "(shared memory)" : path.c_str());
        },
        [&](const std::string& key, const ItemToDecode& it) { return jt9.run(key, it); });
    pool->init();

    // one receiver per 192 kHz band segment (findBand(), source/CWSL_Utils.hpp:27-53, picks the CWSL
    // band containing the frequency; here bands are synthetic: LO = freq rounded to 100 kHz)
    // signal time is simulated per receiver, so each receiver gets its own set of slot clocks
    // (the reference has one wall clock for all, source/CWSL_DIGI.cpp:1134-1175)
    std::map<FrequencyHz, std::shared_ptr<SyncPredicates>> preds;
    std::map<FrequencyHz, std::shared_ptr<Receiver>> receivers;
    std::map<FrequencyHz, std::set<float>> periods;
    std::map<FrequencyHz, std::vector<SyntheticIqSource::Carrier>> carriers;
    std::map<FrequencyHz, std::vector<FskBurst>> bursts;
    float longest = 0;
    for (auto& d : cfg.decoders) {
        const FrequencyHz lo = (d.getFreqCalibrated() + 50000) / 100000 * 100000;
        periods[lo].insert(d.getTRPeriod());
        longest = std::max(longest, d.getTRPeriod());
        if (d.getMode() == "WSPR") {
            // WSPR decoders hear three valid transmissions per 120 s slot (WsprSynth.hpp) instead of a plain carrier:
            // what reaches wsprd through the WAV hand-off is decodable (tests/test_host_gpu.py decodes it)
            for (const auto& tx : kWsprDemoTransmissions) {
                FskBurst b;
                std::string why;
                if (!makeWsprBurst(tx.call, tx.grid, tx.dbm, d.getFreqCalibrated() + tx.audio_hz,
                                   amplitudeForSnr(tx.snr_db, 300.0, 192000.0), b, tx.t0_s, &why)) {
                    printer->err("WSPR demo transmission: " + why);
                    return EXIT_FAILURE;
                }
                bursts[lo].push_back(b);
            }
            continue;
        }
        carriers[lo].push_back({d.getFreqCalibrated() + 1500.0, 8000.0});
    }
    const std::uint32_t fs = 192000, iq_len = 2048;
    const std::uint64_t max_blocks = static_cast<std::uint64_t>(slots * longest * fs / iq_len) + 2;
    int ridx = 0;
    std::vector<std::unique_ptr<ShmProducer>> producers;
    for (auto& kv : periods) {
        preds[kv.first] = std::make_shared<SyncPredicates>();
        auto synth = std::make_unique<SyntheticIqSource>(fs, iq_len, kv.first, carriers[kv.first], 300.0, 20261017 + ridx, max_blocks);
        for (const FskBurst& b : bursts[kv.first]) synth->addBurst(b);
        std::unique_ptr<IqSource> inner;
        std::string smname = "SYNTH" + std::to_string(kv.first / 1000) + "kHz";
        if (shmSource) {
            smname = createSharedMemName(ridx, -1) + "_demo" + std::to_string(::getpid());
            producers.push_back(std::make_unique<ShmProducer>(smname, std::move(synth), 256));
            if (!producers.back()->start()) {
                printer->err("cannot create shared memory segment " + smname);
                return EXIT_FAILURE;
            }
            inner = std::make_unique<CountingShmSource>(producers.back().get());
        } else {
            inner = std::move(synth);
        }
        auto src = std::make_unique<ClockedSource>(std::move(inner), preds[kv.first], kv.second, dumpDir);
        auto r = std::make_shared<Receiver>(smname, printer, std::move(src), ridx % cwsl_device_count(), mode);
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
    for (auto& p : producers) p->join();
    pool->drain();
    pool->terminate();
    jt9.joinAll();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::uint64_t blocks = 0;
    for (auto& kv : receivers) blocks += kv.second->blocksRead();
    std::printf("station_demo: %zu receivers, %zu decoders, %llu IQ blocks (%.1f s of signal per receiver) in %.2f s wall; "
                "%zu audio buffers handed to the decoder pool, %zu of them through jt9 shared memory (%d handshakes completed)\n",
                receivers.size(), cfg.decoders.size(), (unsigned long long)blocks, slots * longest, sec, handled,
                pool->handledViaShMem(), jt9.terminated.load());
    return handled > 0 ? EXIT_SUCCESS : EXIT_FAILURE;
}
