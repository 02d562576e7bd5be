// Downstream hand-off: DecoderPool::push(const ItemToDecode&) (source/DecoderPool.hpp:174-210,
// :1169-1171). The reference's workers then either write a WAV file or fill jt9's Qt shared
// memory and spawn jt9.exe / wsprd.exe / js8.exe with CreateProcessA (:316-415, :600-1167).
// Process spawning of the external WSJT-X binaries is out of scope here (Win32-only, binaries not
// available); this pool keeps the queue, the age check and the WAV hand-off (row f1), and gives
// the finished WAV path (or the item itself) to a caller-supplied sink.
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "CWSL_DIGI_Types.hpp"
#include "WaveFile.hpp"

struct ItemToDecode {  // source/DecoderPool.hpp:174-210
    std::string mode = "";
    std::uint64_t epochTime = 0;   // slot start, seconds since the epoch (source/Instance.cpp:215)
    FrequencyHz baseFreq = 0;      // un-calibrated dial frequency (ssbFreq)
    std::vector<std::int16_t> audio;
    int instanceId = 0;
    std::string cwd;
    float trperiod = 0;

    ItemToDecode() = default;
    ItemToDecode(std::vector<std::int16_t> audioIn, const std::string modeIn, const std::uint64_t epochTimeIn,
                 const FrequencyHz baseFreqIn, const int instanceIdIn, const std::string& cwdIn,
                 const float trperiodIn)
        : mode(modeIn), epochTime(epochTimeIn), baseFreq(baseFreqIn), audio(std::move(audioIn)),
          instanceId(instanceIdIn), cwd(cwdIn), trperiod(trperiodIn) {}
};

class DecoderPool {
public:
    // sink(item, wavPath): called on a worker thread once the hand-off artefact exists; wavPath is
    // empty when transferMethod is not "wavefile".
    using Sink = std::function<void(const ItemToDecode&, const std::string&)>;

    DecoderPool(const std::string& transferMethodIn, const std::string& wavPathIn, int nWorkers,
                std::uint64_t maxDataAgeSec, std::shared_ptr<ScreenPrinter> sp, Sink sinkIn = nullptr)
        : transferMethod(transferMethodIn), wavPath(wavPathIn), numWorkers(nWorkers), maxDataAge(maxDataAgeSec),
          screenPrinter(std::move(sp)), sink(std::move(sinkIn)) {}
    ~DecoderPool() { terminate(); }

    bool init() {  // source/DecoderPool.hpp:252-266
        terminateFlag = false;
        for (int i = 0; i < numWorkers; ++i) workers.emplace_back(&DecoderPool::doWork, this);
        return true;
    }

    // thread-safe; the audio is copied into the queue like the reference does (by-value item)
    void push(const ItemToDecode& item) {
        {
            std::lock_guard<std::mutex> lk(mu);
            queue.push_back(item);
        }
        cv.notify_one();
    }
    void push(ItemToDecode&& item) {
        {
            std::lock_guard<std::mutex> lk(mu);
            queue.push_back(std::move(item));
        }
        cv.notify_one();
    }

    void terminate() {
        terminateFlag = true;
        cv.notify_all();
        for (auto& t : workers)
            if (t.joinable()) t.join();
        workers.clear();
    }

    // block until the queue is drained and no worker is busy (demo/tests)
    void drain() {
        std::unique_lock<std::mutex> lk(mu);
        idle.wait(lk, [&] { return queue.empty() && busy == 0; });
    }

    std::size_t handled() const { return nHandled.load(); }
    std::size_t droppedForAge() const { return nDropped.load(); }

private:
    void doWork() {  // source/DecoderPool.hpp:316-415 (dequeue, age check, hand-off)
        for (;;) {
            ItemToDecode item;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait_for(lk, std::chrono::milliseconds(250), [&] { return terminateFlag || !queue.empty(); });
                if (queue.empty()) {
                    if (terminateFlag) return;
                    continue;
                }
                item = std::move(queue.front());
                queue.pop_front();
                ++busy;
            }
            const std::uint64_t now = std::chrono::system_clock::now().time_since_epoch() / std::chrono::seconds(1);
            const std::uint64_t age = now > item.epochTime ? now - item.epochTime : 0;  // :357-377
            if (maxDataAge && item.epochTime && age > maxDataAge + static_cast<std::uint64_t>(item.trperiod)) {
                screenPrinter->err("Data too old, skipping decode. Age: " + std::to_string(age) + " sec");
                ++nDropped;
            } else {
                std::string path;
                if (transferMethod == "wavefile") {
                    path = wavPath + "/" + std::to_string(item.epochTime) + "_" + std::to_string(item.baseFreq) + "_" +
                           item.mode + "_" + std::to_string(item.instanceId) + "_" + std::to_string(nFiles++) + ".wav";
                    if (!waveWrite(item.audio, path)) {
                        screenPrinter->err("Error writing wave file data: " + path);
                        path.clear();
                    }
                }
                if (sink) sink(item, path);
                ++nHandled;
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                --busy;
            }
            idle.notify_all();
        }
    }

    std::string transferMethod, wavPath;
    int numWorkers;
    std::uint64_t maxDataAge;
    std::shared_ptr<ScreenPrinter> screenPrinter;
    Sink sink;
    std::mutex mu;
    std::condition_variable cv, idle;
    std::deque<ItemToDecode> queue;
    std::vector<std::thread> workers;
    std::atomic_bool terminateFlag{false};
    int busy = 0;
    std::atomic<std::size_t> nHandled{0}, nDropped{0}, nFiles{0};  // nFiles: the reference names files by uuid (:901)
};
