// Downstream hand-off: DecoderPool::push(const ItemToDecode&) (source/DecoderPool.hpp:174-210,
// :1169-1171). The reference's workers then either write a WAV file or fill jt9's Qt shared
// memory and spawn jt9.exe / wsprd.exe / js8.exe with CreateProcessA (:316-415, :600-1167).
// Process spawning of the external WSJT-X binaries is out of scope here (Win32-only, binaries not
// available); this pool keeps the queue, the age check and BOTH hand-offs -- the WAV file (row f1,
// transfermethod=wavefile) and jt9's shared-memory block with its ipc[] handshake (row f2,
// transfermethod=shmem, source/DecoderPool.hpp:379-395, :421-593, :689-709) -- and gives the finished
// artefact to caller-supplied hooks: `sink` (WAV path or item) and `shmDecoder` (the stand-in for
// CreateProcessA("jt9 ... -s <key>") + reading its output).
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
#include "ItemToDecode.hpp"
#include "Jt9SharedMemory.hpp"
#include "WaveFile.hpp"

class DecoderPool {
public:
    // sink(item, wavPath): called on a worker thread once the hand-off artefact exists; wavPath is
    // empty when transferMethod is not "wavefile".
    using Sink = std::function<void(const ItemToDecode&, const std::string&)>;
    // shmDecoder(key, item): called on a worker thread once the segment `key` holds the filled dec_data_t. Stands
    // in for starting "jt9 <mode flags> -s <key>" and reading its output to the end
    // (source/DecoderPool.hpp:634-687); returns what readDataFromExtProgram returns (true = the decoder ran and
    // printed its result). The decoder itself keeps the segment attached until it sees ipc[2] == 1.
    using ShmDecoder = std::function<bool(const std::string& key, const ItemToDecode&)>;

    DecoderPool(const std::string& transferMethodIn, const std::string& wavPathIn, int nWorkers,
                std::uint64_t maxDataAgeSec, std::shared_ptr<ScreenPrinter> sp, Sink sinkIn = nullptr,
                ShmDecoder shmDecoderIn = nullptr, int highestDecodeFreqIn = 3000, int decodeDepthIn = 3)
        : transferMethod(transferMethodIn), wavPath(wavPathIn), numWorkers(nWorkers), maxDataAge(maxDataAgeSec),
          screenPrinter(std::move(sp)), sink(std::move(sinkIn)), shmDecoder(std::move(shmDecoderIn)),
          highestDecodeFreq(highestDecodeFreqIn), decodedepth(decodeDepthIn) {}
    ~DecoderPool() { terminate(); }

    bool init() {  // source/DecoderPool.hpp:252-266
        terminateFlag = false;
        for (int i = 0; i < numWorkers; ++i) workers.emplace_back(&DecoderPool::doWork, this, static_cast<std::size_t>(i));
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
    std::size_t handledViaShMem() const { return nShm.load(); }
    std::size_t droppedForAge() const { return nDropped.load(); }

private:
    // source/DecoderPool.hpp:421-593 + :689-709: segment of sizeof(dec_data_t), per-mode parameter fill, audio into
    // d2[], decoder run, then the ipc[] handshake -- wait until the decoder has cleared ipc[1], answer
    // ipc[1] = 999 / ipc[2] = 1 (terminate), detach.
    bool decodeUsingShMem(const ItemToDecode& item, const std::size_t workerIndex) {
        const std::uint64_t ms = std::chrono::duration_cast<std::chrono::milliseconds>(
                                     std::chrono::system_clock::now().time_since_epoch()).count();
        const std::string skey = "CWSL_DIGI_" + std::to_string(workerIndex) + "_" + std::to_string(item.instanceId) + "_" +
                                 std::to_string(ms) + "_" + std::to_string(nFiles++);
        Jt9ShmSegment mem_jt9;
        if (!mem_jt9.create(skey)) {
            screenPrinter->err("Failed to create shared memory segment! key=" + skey);
            return false;
        }
        dec_data_t* dec_data = mem_jt9.data();
        if (!fillDecData(dec_data, item, highestDecodeFreq, decodedepth)) {
            screenPrinter->err("Unknown mode : " + item.mode);
            return false;
        }
        std::atomic_thread_fence(std::memory_order_seq_cst);  // mem_jt9.unlock(): the block is complete before the decoder starts
        const bool extStatus = shmDecoder ? shmDecoder(skey, item) : false;
        if (extStatus) {
            volatile int* ipc = dec_data->ipc;
            const auto t0 = std::chrono::steady_clock::now();
            while (ipc[1] != 0) {  // the decoder clears istart when it is done with the data
                if (terminateFlag || std::chrono::steady_clock::now() - t0 > std::chrono::seconds(90)) break;
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
            }
            ipc[1] = 999;
            ipc[2] = 1;  // tell the decoder to terminate
            std::atomic_thread_fence(std::memory_order_seq_cst);
        }
        ++nShm;
        return extStatus;  // ~Jt9ShmSegment: detach + unlink
    }

    void doWork(const std::size_t workerIndex) {  // source/DecoderPool.hpp:316-415 (dequeue, age check, hand-off)
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
                // WSPR, JS8 and the FST4 / FST4W family always go through WAV files; FT8, FT4, JT65 and Q65 use
                // shared memory when transfermethod=shmem (source/DecoderPool.hpp:379-395)
                const bool viaFile = transferMethod != "shmem" || item.mode == "WSPR" || item.mode == "JS8" ||
                                     item.mode.rfind("FST4", 0) == 0;
                if (!viaFile) {
                    decodeUsingShMem(item, workerIndex);
                } else if (transferMethod == "wavefile" || transferMethod == "shmem") {
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
    ShmDecoder shmDecoder;
    int highestDecodeFreq, decodedepth;
    std::mutex mu;
    std::condition_variable cv, idle;
    std::deque<ItemToDecode> queue;
    std::vector<std::thread> workers;
    std::atomic_bool terminateFlag{false};
    int busy = 0;
    std::atomic<std::size_t> nHandled{0}, nDropped{0}, nShm{0}, nFiles{0};  // nFiles: the reference names files by uuid (:901)
};
