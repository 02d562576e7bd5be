// CPU-only unit tests of the host-side classes (no GPU call is made): decoder= grammar,
// calibration arithmetic, mode/period table, WAV header bytes, DecoderPool hand-off, predicates.
// Run by tests/test_host_cpu.py; exit code 0 = all passed.
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <sstream>

#include <future>

#include "cwsl_host.hpp"

static int g_fail = 0;
#define EXPECT(cond)                                                       \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++g_fail;                                                      \
        }                                                                  \
    } while (0)
template <class F>
static bool throws(F f) {
    try {
        f();
    } catch (const std::exception&) {
        return true;
    }
    return false;
}

// "K1ABC FN42 37": the channel symbols printed in the WSPR protocol description (sync vector in the LSBs)
static const char kWsprKat[] =
    "330020001020131222100323133220200032012322002232110233210221321222033030301210212032132003323032203020201023"
    "021112330231212221332000010320132222202332323320031222";

int main(int argc, char** argv) {
    // host_tests wspr CALL GRID DBM            -> the 162 channel symbols (tests/test_wspr_cpu.py compares encoders)
    // host_tests wspriq OUT.f32 SECONDS SNR_DB -> SyntheticIqSource with one K1ABC transmission at LO + 1500 Hz
    if (argc >= 5 && std::string(argv[1]) == "wspr") {
        std::array<std::uint8_t, wspr::kSymbols> s{};
        std::string why;
        if (!wspr::encode(argv[2], argv[3], std::atoi(argv[4]), s, &why)) {
            std::fprintf(stderr, "%s\n", why.c_str());
            return 3;
        }
        for (auto v : s) std::putchar('0' + v);
        std::putchar('\n');
        return 0;
    }
    if (argc >= 5 && std::string(argv[1]) == "wspriq") {
        const std::uint32_t fs = 192000, iq_len = 2048;
        SyntheticIqSource src(fs, iq_len, 14100000, {}, 300.0, 20261017);
        FskBurst b;
        if (!makeWsprBurst("K1ABC", "FN42", 37, 14100000 - 4400 + 1500.0, amplitudeForSnr(std::atof(argv[4]), 300.0, fs), b)) return 3;
        src.addBurst(b);
        FILE* f = std::fopen(argv[2], "wb");
        if (!f) return 3;
        std::vector<float> blk(2 * iq_len);
        const std::uint64_t nblk = static_cast<std::uint64_t>(std::atof(argv[3]) * fs / iq_len);
        for (std::uint64_t i = 0; i < nblk && src.readBlock(blk.data()); ++i) std::fwrite(blk.data(), sizeof(float), blk.size(), f);
        std::fclose(f);
        return 0;
    }
    const std::string tmp = argc > 1 ? argv[1] : "/tmp";

    // ---- WSPR channel code (WsprSynth.hpp): known-answer test + rejections ----
    {
        std::array<std::uint8_t, wspr::kSymbols> s{};
        EXPECT(wspr::encode("K1ABC", "FN42", 37, s));
        bool same = true;
        for (int i = 0; i < wspr::kSymbols; ++i) same = same && s[i] == kWsprKat[i] - '0';
        EXPECT(same);
        std::array<std::uint8_t, wspr::kSymbols> s2{};
        EXPECT(wspr::encode(" k1abc ", "fn42", 37, s2) && s2 == s);           // case and padding do not matter
        std::string why;
        EXPECT(!wspr::encode("N0CALL", "FN20", 30, s2, &why) && !why.empty()); // digit would have to sit in 4th place
        EXPECT(!wspr::encode("K1ABC", "FN4", 37, s2) && !wspr::encode("K1ABC", "ZZ42", 37, s2));
        EXPECT(!wspr::encode("K1ABC", "FN42", 35, s2));                        // not a WSPR power level
        FskBurst b;
        EXPECT(makeWsprBurst("W1AW", "FN31", 30, 14097100.0, 10.0, b) && b.symbols.size() == 162 && b.period_s == 120.0);
        EXPECT(std::fabs(b.rf_hz - (14097100.0 - 1.5 * 12000.0 / 8192.0)) < 1e-9);
        EXPECT(std::fabs(amplitudeForSnr(0.0, 300.0, 192000.0) - std::sqrt(2.0 * 90000.0 * 2500.0 / 192000.0)) < 1e-9);
    }

    // ---- periods (source/CWSL_DIGI.hpp:64-113) ----
    EXPECT(getRXPeriod("FT8") == 15.0f && getRXPeriod("FT4") == 7.5f && getRXPeriod("WSPR") == 120.0f);
    EXPECT(getRXPeriod("JT65") == 60.0f && getRXPeriod("JS8") == 15.0f && getRXPeriod("Q65-30") == 30.0f);
    EXPECT(getRXPeriod("FST4W-1800") == 1800.0f && getRXPeriod("FST4-300") == 300.0f);
    EXPECT(throws([] { getRXPeriod("PSK31"); }));

    // ---- decoder= grammar (source/CWSL_DIGI.cpp:731-842) ----
    {
        Decoder d = parseDecoderLine("14074000 FT8", 1.0, -1, "W1AW");
        EXPECT(d.getFreq() == 14074000u && d.getFreqCalibrated() == 14074000u && d.getMode() == "FT8");
        EXPECT(d.getsmNum() == -1 && d.getReporterCallsign() == "W1AW" && d.getTRPeriod() == 15.0f);
        Decoder e = parseDecoderLine("14095600 WSPR 2 1.0000015 K1ABC", 1.00000071, -1, "W1AW");
        EXPECT(e.getsmNum() == 2 && e.getReporterCallsign() == "K1ABC");
        EXPECT(e.getFreqCalibrated() == static_cast<FrequencyHz>(14095600u / (1.00000071 * 1.0000015)));
        EXPECT(e.getFreqCalibrated() == 14095568u);
        EXPECT(throws([] { parseDecoderLine("14074000", 1.0, -1, ""); }));                    // 1 token
        EXPECT(throws([] { parseDecoderLine("14074000 FT8 0 1.0 W1AW X", 1.0, -1, ""); }));    // 6 tokens
        EXPECT(throws([] { parseDecoderLine("14074000 PSK31", 1.0, -1, ""); }));              // unknown mode
        EXPECT(throws([] { parseDecoderLine("14074000 FT8 0 1.0 W1AW", 1.0, -1, ""); }));      // callsign on non-WSPR
        EXPECT(throws([] { parseDecoderLine("14074000  FT8", 1.0, -1, ""); }));  // double space -> empty mode token
    }
    {
        std::istringstream ini(
            "[radio]\nfreqcalibration=1.0\nsharedmem=-1\n[operator]\ncallsign=N0CALL\n"
            "[decoders]\n# 20m\ndecoder=14095600 WSPR\ndecoder=14074000 FT8\ndecoder=14080000 FT4\n"
            "[wsjtx]\nftaudioscalefactor=0.85\n");
        FrontEndConfig c = loadFrontEndConfig(ini);
        EXPECT(c.decoders.size() == 3 && c.ftAudioScaleFactor == 0.85f && c.wsprAudioScaleFactor == 0.20f);
        EXPECT(c.decoders[0].getReporterCallsign() == "N0CALL");
        std::istringstream bad("[decoders]\ndecoder=14074000 FT8\n[wsjtx]\nftaudioscalefactor=1.5\n");
        EXPECT(throws([&] { loadFrontEndConfig(bad); }));
        EXPECT(c.kernelMode == CWSL_MODE_FAST && c.cudaDevice == 0);
        std::istringstream gpu("[decoders]\ndecoder=14074000 FT8\n[gpu]\narithmetic=stft\ndevice=3\n");
        FrontEndConfig cg = loadFrontEndConfig(gpu);
        EXPECT(cg.kernelMode == CWSL_MODE_STFT && cg.cudaDevice == 3);
        std::istringstream gbad("[decoders]\ndecoder=14074000 FT8\n[gpu]\narithmetic=double\n");
        EXPECT(throws([&] { loadFrontEndConfig(gbad); }));
        std::istringstream none("[radio]\nfreqcalibration=1.0\n");
        EXPECT(throws([&] { loadFrontEndConfig(none); }));
    }

    // ---- WAV hand-off (source/WaveFile.hpp:19-35; DecoderPool.hpp:915-964) ----
    {
        std::vector<std::int16_t> audio(240000);
        for (size_t i = 0; i < audio.size(); ++i) audio[i] = static_cast<std::int16_t>(i * 7 - 1000);
        const std::string path = tmp + "/cwsl_host_test.wav";
        EXPECT(waveWrite(audio, path));
        FILE* f = std::fopen(path.c_str(), "rb");
        EXPECT(f != nullptr);
        if (f) {
            unsigned char h[46];
            EXPECT(std::fread(h, 1, 46, f) == 46);
            EXPECT(std::memcmp(h, "RIFF", 4) == 0 && std::memcmp(h + 8, "WAVEfmt ", 8) == 0);
            auto u32 = [&](int o) { return h[o] | (h[o + 1] << 8) | (h[o + 2] << 16) | ((unsigned)h[o + 3] << 24); };
            auto u16 = [&](int o) { return h[o] | (h[o + 1] << 8); };
            EXPECT(u32(4) == 46 + 480000 - 8 && u32(16) == 18 && u16(20) == 1 && u16(22) == 1);
            EXPECT(u32(24) == 12000 && u32(28) == 24000 && u16(32) == 2 && u16(34) == 16 && u16(36) == 0);
            EXPECT(std::memcmp(h + 38, "data", 4) == 0 && u32(42) == 480000);
            std::fseek(f, 0, SEEK_END);
            EXPECT(std::ftell(f) == 46 + 480000);
            std::int16_t s[2];
            std::fseek(f, 46 + 2 * 1000, SEEK_SET);
            EXPECT(std::fread(s, 2, 2, f) == 2 && s[0] == audio[1000] && s[1] == audio[1001]);
            std::fclose(f);
        }
        std::remove(path.c_str());
    }

    // ---- DecoderPool: push -> worker -> WAV + sink; age check (source/DecoderPool.hpp:357-377) ----
    {
        auto sp = std::make_shared<ScreenPrinter>(LOG_LEVEL::ERR);
        std::vector<std::string> seen;
        std::mutex mu;
        auto pool = std::make_shared<DecoderPool>("wavefile", tmp, 2, 300, sp,
                                                  [&](const ItemToDecode& it, const std::string& path) {
                                                      std::lock_guard<std::mutex> lk(mu);
                                                      seen.push_back(it.mode + ":" + path);
                                                      if (!path.empty()) std::remove(path.c_str());
                                                  });
        pool->init();
        const std::uint64_t now = std::chrono::system_clock::now().time_since_epoch() / std::chrono::seconds(1);
        for (int i = 0; i < 5; ++i)
            pool->push(ItemToDecode(std::vector<std::int16_t>(150000, (std::int16_t)i), "FT4", now, 14080000, i, "cwd", 7.5f));
        pool->push(ItemToDecode(std::vector<std::int16_t>(240000), "FT8", now - 10000, 14074000, 9, "cwd", 15.0f));  // stale
        pool->drain();
        pool->terminate();
        EXPECT(seen.size() == 5 && pool->handled() == 5 && pool->droppedForAge() == 1);
        for (auto& s : seen) EXPECT(s.rfind("FT4:", 0) == 0 && s.find(".wav") != std::string::npos);
    }

    // ---- predicates (source/CWSL_DIGI_Types.hpp:65-145) ----
    {
        SyncPredicates preds;
        auto a = preds.createPredicate("FT8"), b = preds.createPredicate("FT8"), c = preds.createPredicate("JS8"),
             d = preds.createPredicate("FT4"), w = preds.createPredicate("WSPR"), f = preds.createPredicate("FST4W-120");
        EXPECT(a != b && !a->load());  // one predicate per decoder
        preds.fire(15.0f);  // the 15 s clock thread fires FT8 and JS8 (ft8Preds, source/CWSL_DIGI.cpp:234-262)
        EXPECT(a->load() && b->load() && c->load() && !d->load() && !w->load());
        a->store(false);
        EXPECT(b->load());
        preds.fire(120.0f);
        EXPECT(w->load() && f->load() && !d->load());
        EXPECT(throws([&] { preds.createPredicate("RTTY"); }));
    }

    // ---- Instance frequency plumbing without a GPU (source/Instance.cpp:183, :320-329) ----
    {
        auto sp = std::make_shared<ScreenPrinter>(LOG_LEVEL::ERR);
        auto src = std::make_unique<SyntheticIqSource>(192000, 2048, 14100000, std::vector<SyntheticIqSource::Carrier>{});
        auto rcv = std::make_shared<Receiver>("CWSL20Band0", sp, std::move(src));
        // Receiver::init() needs a GPU; only the source side is exercised here
        float blk[4096];
        SyntheticIqSource s2(192000, 2048, 14100000, {{14075500.0, 8000.0}}, 300.0, 1, 2);
        EXPECT(s2.readBlock(blk) && s2.readBlock(blk) && !s2.readBlock(blk));
        auto pool = std::make_shared<DecoderPool>("none", tmp, 0, 300, sp);
        auto pred = std::make_shared<SyncPredicate>();
        Instance w(rcv, 0, pred, 14095600, 14095600, "WSPR", "N0CALL", 12000, 0.9f, 0.2f, sp, pool, 120.0f);
        Instance f(rcv, 1, pred, 14097000, 14097000, "FST4W-120", "N0CALL", 12000, 0.9f, 0.2f, sp, pool, 120.0f);
        EXPECT(w.audioScale() == 0.2f && f.audioScale() == 0.9f);
        EXPECT(w.getStatus() == InstanceStatus::NOT_INITIALIZED);
    }

    // ---- wall-clock slot edges (source/CWSL_DIGI.cpp:174-451), with an injected clock ----
    {
        auto T = [](int m, int s, int ms) { UtcTime t; t.minute = m; t.second = s; t.millis = ms; return t; };
        EXPECT(slotEdgeNow(15.0f, T(3, 15, 10)) && !slotEdgeNow(15.0f, T(3, 16, 0)) && slotEdgeNow(15.0f, T(3, 0, 999)));
        EXPECT(slotEdgeNow(7.5f, T(0, 30, 0)) && !slotEdgeNow(7.5f, T(0, 7, 299)) && slotEdgeNow(7.5f, T(0, 7, 300)));
        EXPECT(slotEdgeNow(7.5f, T(0, 52, 400)) && !slotEdgeNow(7.5f, T(0, 8, 0)));
        EXPECT(slotEdgeNow(30.0f, T(9, 30, 0)) && !slotEdgeNow(30.0f, T(9, 15, 0)));
        EXPECT(slotEdgeNow(60.0f, T(9, 0, 0)) && !slotEdgeNow(60.0f, T(9, 30, 0)));
        EXPECT(slotEdgeNow(120.0f, T(10, 0, 0)) && !slotEdgeNow(120.0f, T(11, 0, 0)));
        EXPECT(slotEdgeNow(300.0f, T(25, 0, 0)) && !slotEdgeNow(300.0f, T(26, 0, 0)));
        EXPECT(slotEdgeNow(900.0f, T(45, 0, 0)) && !slotEdgeNow(900.0f, T(50, 0, 0)));
        EXPECT(slotEdgeNow(1800.0f, T(30, 0, 0)) && !slotEdgeNow(1800.0f, T(15, 0, 0)));
        // simulated two minutes polled every 25 ms: count the edges each period produces
        auto preds = std::make_shared<SyncPredicates>();
        auto p8 = preds->createPredicate("FT8"), p4 = preds->createPredicate("FT4"), pw = preds->createPredicate("WSPR"),
             pj = preds->createPredicate("JT65");
        std::uint64_t fake = 1792214520000ull;  // an even-minute boundary (divisible by 120000)
        EXPECT(fake % 120000 == 0);
        SlotClocks clk(preds, [&] { return fake; });
        int l8 = -1, l4 = -1, lw = -1, lj = -1, n8 = 0, n4 = 0, nw = 0, nj = 0;
        for (int i = 0; i < 120000 / 25; ++i, fake += 25) {
            n8 += clk.poll(15.0f, l8);
            n4 += clk.poll(7.5f, l4);
            nw += clk.poll(120.0f, lw);
            nj += clk.poll(60.0f, lj);
        }
        EXPECT(n8 == 8 && n4 == 16 && nw == 1 && nj == 2);
        EXPECT(p8->load() && p4->load() && pw->load() && pj->load());
    }

    // ---- CWSL shared-memory layout on POSIX shm (source/SharedMemory.cpp:115-246) ----
    {
        const std::string name = createSharedMemName(7, -1) + "_t" + std::to_string(::getpid());
        EXPECT(createSharedMemName(7, -1) == "CWSL7Band" && createSharedMemName(3, 2) == "CWSL3Band2");
        CSharedMemory wr;
        const std::uint32_t iq_len = 512, blk_bytes = iq_len * 8;
        SM_HDR hdr{192000, (int)iq_len, 14100000};
        EXPECT(wr.Create(name, blk_bytes * 5 + 24, hdr));   // ring not a multiple of the block: wrap splits a block
        CwslShmSource src(200);
        EXPECT(src.open(name));
        EXPECT(src.sampleRate() == 192000 && src.blockInSamples() == iq_len && src.L0() == 14100000u);
        std::vector<float> blk(iq_len * 2), got(iq_len * 2);
        EXPECT(!src.readBlock(got.data()));                  // nothing written yet -> timeout
        bool all = true;
        for (int b = 0; b < 23; ++b) {
            for (std::uint32_t i = 0; i < iq_len * 2; ++i) blk[i] = static_cast<float>(b * 10000 + (int)i);
            EXPECT(wr.Write(reinterpret_cast<const std::uint8_t*>(blk.data()), blk_bytes));
            EXPECT(src.readBlock(got.data()));
            all = all && std::memcmp(blk.data(), got.data(), blk_bytes) == 0;
        }
        EXPECT(all);
        CwslShmSource missing(50);
        EXPECT(!missing.open("CWSL_no_such_band"));
    }

    // ---- jt9 shared-memory block (source/DecoderPool.hpp:44-108, :421-593) ----
    {
        std::unique_ptr<dec_data_t> dd(new dec_data_t);
        std::vector<std::int16_t> audio(240000);
        for (size_t i = 0; i < audio.size(); ++i) audio[i] = static_cast<std::int16_t>(i % 1000 - 500);
        ItemToDecode ft8(audio, "FT8", 1792214520, 14074000, 3, "cwd", 15.0f);
        EXPECT(fillDecData(dd.get(), ft8, 6000, 3));
        EXPECT(dd->params.nmode == 8 && dd->params.ntrperiod == 15 && dd->params.lft8apon && dd->params.napwid == 50);
        EXPECT(dd->params.nfa == 0 && dd->params.nfb == 6000 && dd->params.ndepth == 3 && dd->params.newdat && dd->params.dttol == 4.0f);
        EXPECT(dd->ipc[0] == 0 && dd->ipc[1] == 1 && dd->ipc[2] == -1);
        EXPECT(dd->d2[0] == audio[0] && dd->d2[239999] == audio[239999] && dd->d2[240000] == 0);
        ItemToDecode ft4(std::vector<std::int16_t>(150000, 7), "FT4", 0, 14080000, 4, "cwd", 7.5f);
        EXPECT(fillDecData(dd.get(), ft4, 3000, 2) && dd->params.nmode == 5 && dd->params.ntrperiod == 7 && dd->params.napwid == 80);
        EXPECT(dd->d2[149999] == 7 && dd->d2[150000] == 0);   // block is cleared first
        ItemToDecode f4w(std::vector<std::int16_t>(21660000, 1), "FST4W-1800", 0, 474200, 5, "cwd", 1800.0f);
        EXPECT(fillDecData(dd.get(), f4w, 3000, 3) && dd->params.nmode == 241 && dd->params.nzhsym == 6232 && dd->ipc[0] == 6232);
        EXPECT(dd->params.nfqso == 1500 && dd->params.nexp_decode == 768 && dd->d2[NTMAX * RX_SAMPLE_RATE - 1] == 1);  // clipped at NTMAX
        ItemToDecode f300(audio, "FST4-300", 0, 474200, 6, "cwd", 300.0f);
        EXPECT(fillDecData(dd.get(), f300, 3000, 3) && dd->params.nfa == 700 && dd->params.nfb == 1100 && dd->params.ndepth == 1 && dd->params.nmode == 240);
        ItemToDecode wspr(audio, "WSPR", 0, 14095600, 7, "cwd", 120.0f);
        EXPECT(!fillDecData(dd.get(), wspr, 3000, 3));           // WSPR always goes through a WAV file
    }

    // ---- transfermethod=shmem through the pool: segment life cycle + ipc[] handshake (source/DecoderPool.hpp:421-448,
    //      :575-577, :689-709); WSPR / JS8 / FST4(W) still go through WAV files (:379-395) ----
    {
        auto printer = std::make_shared<ScreenPrinter>(LOG_LEVEL::ERR);
        std::mutex m;
        std::vector<std::string> keys, wavs;
        std::vector<std::thread> decoders;
        std::atomic<int> saw_data{0}, told_to_quit{0};
        const std::string dir = "/tmp";
        {
            DecoderPool pool(
                "shmem", dir, 2, 0, printer,
                [&](const ItemToDecode&, const std::string& path) {
                    std::lock_guard<std::mutex> lk(m);
                    if (!path.empty()) wavs.push_back(path);
                },
                [&](const std::string& key, const ItemToDecode& item) {
                    std::promise<bool> ready;
                    auto fut = ready.get_future();
                    std::lock_guard<std::mutex> lk(m);
                    keys.push_back(key);
                    decoders.emplace_back([&, key, item, p = std::move(ready)]() mutable {  // "jt9 -s <key>"
                        Jt9ShmSegment seg;
                        if (!seg.attach(key)) {
                            p.set_value(false);
                            return;
                        }
                        dec_data_t* d = seg.data();
                        const bool same = std::memcmp(d->d2, item.audio.data(), item.audio.size() * 2) == 0;
                        saw_data += same && d->ipc[1] == 1 && d->ipc[2] == -1 && d->params.nmode == (item.mode == "FT8" ? 8 : 5);
                        p.set_value(true);
                        volatile int* ipc = d->ipc;
                        ipc[1] = 0;
                        for (int i = 0; i < 5000 && ipc[2] != 1; ++i) std::this_thread::sleep_for(std::chrono::milliseconds(1));
                        told_to_quit += ipc[2] == 1 && ipc[1] == 999;
                    });
                    return fut.get();
                });
            pool.init();
            std::vector<std::int16_t> a(240000), b(150000), w(1500000, 3);
            for (size_t i = 0; i < a.size(); ++i) a[i] = static_cast<std::int16_t>(i * 7);
            for (size_t i = 0; i < b.size(); ++i) b[i] = static_cast<std::int16_t>(i * 3 + 1);
            pool.push(ItemToDecode(a, "FT8", 0, 14074000, 1, "cwd", 15.0f));
            pool.push(ItemToDecode(b, "FT4", 0, 14080000, 2, "cwd", 7.5f));
            pool.push(ItemToDecode(w, "WSPR", 0, 14095600, 3, "cwd", 120.0f));
            pool.push(ItemToDecode(a, "FST4W-120", 0, 474200, 4, "cwd", 120.0f));
            pool.drain();
            EXPECT(pool.handled() == 4 && pool.handledViaShMem() == 2);
            pool.terminate();
        }
        for (auto& t : decoders) t.join();
        EXPECT(keys.size() == 2 && saw_data == 2 && told_to_quit == 2);
        EXPECT(wavs.size() == 2);                                  // WSPR and FST4W-120 went through files
        for (const auto& k : keys) {
            Jt9ShmSegment gone;
            EXPECT(!gone.attach(k));                               // detached and unlinked after the handshake
            EXPECT(k.rfind("CWSL_DIGI_", 0) == 0);
        }
        for (const auto& p : wavs) std::remove(p.c_str());
    }

    if (g_fail) {
        std::fprintf(stderr, "%d host test(s) failed\n", g_fail);
        return 1;
    }
    std::printf("host tests ok\n");
    return 0;
}
