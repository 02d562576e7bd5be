// One receiver per CWSL band: interface of source/Receiver.hpp:52-302, GPU front-end underneath.
//
// What changes against the reference: the reference's Receiver only copies IQ blocks into a host
// ring and every Instance (decoder) thread pops each block and runs its own SSBD on it
// (source/Instance.cpp:259-277). Here the Receiver's reader thread pushes each block ONCE into the
// device ring (cwsl_rx_push_iq) and, when a slot group's SyncPredicate has fired, ends the slot
// for ALL decoders of that group in one batched GPU pass (cwsl_rx_end_slot), then builds one
// ItemToDecode per decoder and hands it to DecoderPool::push -- the exact hand-off of
// source/Instance.cpp:238-245. getIQBuffer() therefore has no consumers and is not provided.
//
// Ingest (source/Receiver.hpp:209-276): blocks are read from the IqSource straight into
// a ring of PINNED staging buffers (cwsl_host_alloc) and pushed asynchronously; the reader only ever waits for the
// host-to-device copy that last read the staging buffer it is about to refill (cwsl_rx_push_fence /
// cwsl_rx_wait_fence), never for kernels or for the hand-off copies of finished slots.
//
// Decoders come and go while the receiver runs (the reference's main loop restarts FINISHED instances,
// source/CWSL_DIGI.cpp:1217-1226): addInstance()/removeInstance() on a RUNNING receiver queue the request, and the
// reader thread applies it at the slot group's next edge, where the GPU side changes its channel set as well.
#pragma once

#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cwsl_b200.h"
#include "CWSL_DIGI_Types.hpp"
#include "DecoderPool.hpp"
#include "IqSource.hpp"

class Instance;

class Receiver {
public:
    Receiver(const std::string& smnameIn, std::shared_ptr<ScreenPrinter> sp, std::unique_ptr<IqSource> src,
             int cudaDevice = 0, int mode = CWSL_MODE_FAST)
        : smname(smnameIn), screenPrinter(std::move(sp)), source(std::move(src)), device(cudaDevice), kernelMode(mode) {}
    virtual ~Receiver() {
        if (status != ReceiverStatus::FINISHED) terminate();
        if (rx) cwsl_rx_destroy(rx);
        for (auto& g : groups) cwsl_host_free(g.audio);
        cwsl_host_free(staging);
    }

    // opens the producer and creates the GPU front-end (source/Receiver.hpp:115-176 opens the shared
    // memory and starts readIQ; the thread here starts in start(), once every decoder is attached)
    bool init() {
        terminateFlag = false;
        if (!source || !source->open(smname)) {
            screenPrinter->err(receiverLog() + "Can't open shared memory : " + smname);
            return false;
        }
        radioSR = source->sampleRate();
        iq_len = source->blockInSamples();
        lo = source->L0();
        screenPrinter->print(receiverLog() + "\tSample Rate: " + std::to_string(radioSR) +
                             "\tBlock In Samples: " + std::to_string(iq_len) + "\tLO: " + std::to_string(lo) +
                             "\tShared Memory: " + smname);
        rx = cwsl_rx_create(device, radioSR, static_cast<uint32_t>(iq_len), /*ring_seconds=*/3.0);
        if (!rx) {
            screenPrinter->err(receiverLog() + std::string("GPU front-end: ") + cwsl_last_error());
            return false;
        }
        cwsl_rx_set_mode(rx, kernelMode);
        return true;
    }

    std::size_t getIQLength() const { return iq_len; }
    std::uint32_t getSampleRate() const { return radioSR; }
    FrequencyHz getLO() const { return lo; }
    ReceiverStatus getStatus() { return status.load(); }
    const std::string& getName() const { return smname; }

    // Called by Instance::init(): registers the decoder as a channel of the slot group that belongs to its
    // SyncPredicate. On a RUNNING receiver the decoder joins at its group's next slot edge (the group, i.e. a decoder
    // with the same period, must already exist: the set of periods is fixed once the receiver runs).
    bool addInstance(Instance* inst);
    // Called by Instance::terminate(): the decoder leaves at its group's next slot edge (at once if the receiver is
    // not running); its Instance object must stay alive until getStatus() of the instance reports FINISHED.
    bool removeInstance(Instance* inst);

    // starts the reader thread (readIQ, source/Receiver.hpp:209-276)
    bool start() {
        if (!rx || instances.empty()) return false;
        status = ReceiverStatus::RUNNING;
        iqThread = std::thread(&Receiver::readIQ, this);
        return true;
    }

    void terminate() {
        terminateFlag = true;
        if (iqThread.joinable()) iqThread.join();
        status = ReceiverStatus::FINISHED;
    }

    // run the reader loop on the calling thread until the source ends (demo/tests)
    void runToEnd() {
        status = ReceiverStatus::RUNNING;
        readIQ();
    }

    std::uint64_t blocksRead() const { return nBlocks.load(); }
    std::uint64_t slotsFinished() const { return nSlots.load(); }
    std::uint64_t blocksDropped() const { return nDroppedBlocks.load(); }
    static constexpr std::size_t kStagingBlocks = 16;  // pinned IQ staging ring (~170 ms at 192 kHz / 2048)

private:
    // Decoders of this receiver that share a period = one GPU slot group. Every member keeps its own
    // SyncPredicate like in the reference; the clock thread sets them in creation order, so the
    // first member's flag is the group's edge and all members' flags are consumed together.
    struct SlotGroup {
        std::vector<std::shared_ptr<SyncPredicate>> preds;
        int id = -1;  // cwsl group id
        float period = 0;
        std::vector<Instance*> members;
        std::uint64_t startEpochTime = 0;  // 0 = first, partial buffer -> discarded (Instance.cpp:224-227)
        std::int16_t* audio = nullptr;     // [members][af_size], pinned hand-off buffer (cwsl_host_alloc)
        std::size_t audioElems = 0;
        std::vector<Instance*> joining, leaving;  // applied at the next slot edge (guarded by `mu`)
        bool idle = false;  // every decoder of the group has terminated: the GPU group keeps one channel, nothing is handed on
    };

    void readIQ();
    void finishSlot(SlotGroup& g);
    void applyMembership(SlotGroup& g);
    bool attach(SlotGroup& g, Instance* inst);
    std::string receiverLog() const { return "Receiver " + smname + " "; }

    std::string smname;
    std::shared_ptr<ScreenPrinter> screenPrinter;
    std::unique_ptr<IqSource> source;
    int device, kernelMode;
    cwsl_rx_t* rx = nullptr;
    std::uint32_t radioSR = 0;
    std::size_t iq_len = 0;
    FrequencyHz lo = 0;
    std::vector<Instance*> instances;
    std::vector<SlotGroup> groups;
    std::thread iqThread;
    std::atomic<ReceiverStatus> status{ReceiverStatus::NOT_INITIALIZED};
    std::atomic_bool terminateFlag{false};
    std::atomic<std::uint64_t> nBlocks{0}, nSlots{0}, nDroppedBlocks{0};
    std::mutex mu;            // membership requests from other threads
    float* staging = nullptr; // [kStagingBlocks][2 * iq_len], pinned
    std::uint64_t stagingFence[kStagingBlocks] = {};
};
