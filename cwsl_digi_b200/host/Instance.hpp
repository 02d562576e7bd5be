// One per decoder: interface of source/Instance.hpp:47-134. The reference's Instance owns a
// thread, an SSBD<float> and a 2-deep float audio ring; here it is a descriptor: its channel
// lives in the Receiver's GPU slot group, and the Receiver's reader thread performs the
// slot-edge work of sampleManager() (source/Instance.cpp:203-253) for the whole group at once.
#pragma once

#include <atomic>
#include <memory>
#include <string>

#include "CWSL_DIGI_Types.hpp"
#include "DecoderPool.hpp"
#include "Receiver.hpp"

class Instance {
public:
    Instance(std::shared_ptr<Receiver> receiverIn, const size_t idIn, std::shared_ptr<SyncPredicate> predIn,
             const FrequencyHz ssbFreqIn, const FrequencyHz calibratedFreqIn, const std::string& modeIn,
             const std::string& callsignIn, const uint32_t waveSampleRateIn, const float audioScaleFactor_ftIn,
             const float audioScaleFactor_wsprIn, std::shared_ptr<ScreenPrinter> sp, std::shared_ptr<DecoderPool> dp,
             const float trperiodIn)
        : ssbFreq(ssbFreqIn), calibratedSSBFreq(calibratedFreqIn), digitalMode(modeIn), waveSampleRate(waveSampleRateIn),
          screenPrinter(std::move(sp)), audioScaleFactor_ft(audioScaleFactor_ftIn),
          audioScaleFactor_wspr(audioScaleFactor_wsprIn), decoderPool(std::move(dp)), pred(std::move(predIn)), id(idIn),
          receiver(std::move(receiverIn)), callsign(callsignIn), status(InstanceStatus::NOT_INITIALIZED),
          trperiod(trperiodIn) {}
    virtual ~Instance() = default;

    std::string getMode() const { return digitalMode; }
    FrequencyHz getFrequency() const { return ssbFreq; }
    FrequencyHz getCalibratedFrequency() const { return calibratedSSBFreq; }
    std::string getCallsign() const { return callsign; }
    InstanceStatus getStatus() { return status.load(); }
    std::size_t getId() const { return id; }
    float getTRPeriod() const { return trperiod; }
    std::shared_ptr<SyncPredicate> getPredicate() const { return pred; }
    std::shared_ptr<DecoderPool> getDecoderPool() const { return decoderPool; }
    const std::string& getCwd() const { return cwd; }

    // source/Instance.cpp:183: int32(calibratedSSBFreq - LO), unsigned arithmetic wrapping into int32
    std::int32_t demodFreq() const { return static_cast<std::int32_t>(calibratedSSBFreq - receiver->getLO()); }
    // source/Instance.cpp:320-329: WSPR uses the wspr factor, everything else (FST4W too) the ft factor
    float audioScale() const { return digitalMode == "WSPR" ? audioScaleFactor_wspr : audioScaleFactor_ft; }

    // source/Instance.cpp:121-176 sizes buffers and starts the thread; here: attach to the receiver
    bool init() {
        cwd = "instance_" + std::to_string(id);
        if (!receiver || !receiver->addInstance(this)) {
            screenPrinter->err(instanceLog() + "could not attach to the receiver: " + cwsl_last_error());
            return false;
        }
        status = InstanceStatus::RUNNING;
        return true;
    }
    // source/Instance.cpp:100-119: stops the decoder. Its channel leaves the receiver's slot group at the next slot
    // edge; the status turns FINISHED once it has (at once if the receiver is not running).
    void terminate() {
        if (status != InstanceStatus::RUNNING || !receiver || !receiver->removeInstance(this)) status = InstanceStatus::FINISHED;
    }
    void markFinished() { status = InstanceStatus::FINISHED; }
    std::string instanceLog() const { return "Instance " + std::to_string(id) + " "; }

    // set by Receiver::addInstance
    int group = -1, channel = -1;

private:
    FrequencyHz ssbFreq;
    FrequencyHz calibratedSSBFreq;
    std::string digitalMode;
    std::uint32_t waveSampleRate;
    std::shared_ptr<ScreenPrinter> screenPrinter;
    float audioScaleFactor_ft;
    float audioScaleFactor_wspr;
    std::shared_ptr<DecoderPool> decoderPool;
    std::shared_ptr<SyncPredicate> pred;
    std::size_t id;
    std::shared_ptr<Receiver> receiver;
    std::string cwd;
    std::string callsign;
    std::atomic<InstanceStatus> status;
    float trperiod;
};

// ---- Receiver members that need the complete Instance type ------------------------------------------
inline bool Receiver::attach(SlotGroup& g, Instance* inst) {
    const int ch = cwsl_rx_add_channel(rx, g.id, inst->demodFreq(), USB, inst->audioScale());
    if (ch < 0) return false;  // e.g. "Signal outside of band", source/SSBD.hpp:100-103
    inst->group = g.id;
    inst->channel = ch;
    g.members.push_back(inst);
    g.preds.push_back(inst->getPredicate());
    instances.push_back(inst);
    return true;
}

inline bool Receiver::addInstance(Instance* inst) {
    if (!rx) return false;
    std::lock_guard<std::mutex> lk(mu);
    SlotGroup* g = nullptr;
    for (auto& sg : groups)
        if (sg.period == inst->getTRPeriod()) g = &sg;
    if (status == ReceiverStatus::RUNNING) {
        // The reader thread owns the GPU handle while it runs: validate here, join at the group's next slot edge.
        if (!g) return false;
        if (cwsl_build_tables(radioSR, inst->demodFreq(), USB, nullptr, nullptr, nullptr) != CWSL_OK) return false;
        g->joining.push_back(inst);
        return true;
    }
    if (!g) {
        SlotGroup sg;
        sg.period = inst->getTRPeriod();
        sg.id = cwsl_rx_add_group(rx, sg.period);
        if (sg.id < 0) return false;
        groups.push_back(std::move(sg));
        g = &groups.back();
    }
    return attach(*g, inst);
}

inline bool Receiver::removeInstance(Instance* inst) {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& g : groups) {
        for (std::size_t m = 0; m < g.members.size(); ++m) {
            if (g.members[m] != inst) continue;
            if (status == ReceiverStatus::RUNNING) {
                g.leaving.push_back(inst);
                return true;  // finishes at the next slot edge
            }
            return false;     // not running: nothing to wait for, the caller marks itself FINISHED
        }
        for (std::size_t j = 0; j < g.joining.size(); ++j)
            if (g.joining[j] == inst) {
                g.joining.erase(g.joining.begin() + j);
                return false;
            }
    }
    return false;
}

// Slot edge of the group, right after cwsl_rx_end_slot: no IQ of the next slot has been pushed yet, so the GPU side
// applies channel-set changes at once and both sides switch for the same slot.
inline void Receiver::applyMembership(SlotGroup& g) {
    std::vector<Instance*> joining, leaving;
    {
        std::lock_guard<std::mutex> lk(mu);
        joining.swap(g.joining);
        leaving.swap(g.leaving);
    }
    for (Instance* inst : leaving) {
        std::size_t m = 0;
        while (m < g.members.size() && g.members[m] != inst) ++m;
        if (m == g.members.size()) continue;
        if (g.members.size() == 1 && joining.empty()) {
            // the GPU group keeps its last channel (a group cannot be empty); its audio is simply no longer handed on
            screenPrinter->debug(receiverLog() + "last decoder of a slot group terminated; the group idles");
            g.idle = true;
        } else if (cwsl_rx_remove_channel(rx, g.id, static_cast<int>(m)) != CWSL_OK) {
            screenPrinter->err(receiverLog() + std::string("remove decoder: ") + cwsl_last_error());
            continue;
        }
        if (!g.idle) {
            g.members.erase(g.members.begin() + m);
            g.preds.erase(g.preds.begin() + m);
            for (std::size_t k = m; k < g.members.size(); ++k) g.members[k]->channel = static_cast<int>(k);
        }
        for (std::size_t k = 0; k < instances.size(); ++k)
            if (instances[k] == inst) {
                instances.erase(instances.begin() + k);
                break;
            }
        inst->markFinished();
    }
    for (Instance* inst : joining) {
        if (g.idle) {  // the idling channel is replaced: add first, then drop the old one
            Instance* old = g.members.empty() ? nullptr : g.members[0];
            if (!attach(g, inst)) {
                screenPrinter->err(receiverLog() + std::string("add decoder: ") + cwsl_last_error());
                inst->markFinished();
                continue;
            }
            if (old && cwsl_rx_remove_channel(rx, g.id, 0) == CWSL_OK) {
                g.members.erase(g.members.begin());
                g.preds.erase(g.preds.begin());
                for (std::size_t k = 0; k < g.members.size(); ++k) g.members[k]->channel = static_cast<int>(k);
            }
            g.idle = false;
            continue;
        }
        if (!attach(g, inst)) {
            screenPrinter->err(receiverLog() + std::string("add decoder: ") + cwsl_last_error());
            inst->markFinished();
        }
    }
}

inline void Receiver::finishSlot(SlotGroup& g) {
    // source/Instance.cpp:203-253 for every decoder of the group
    const std::uint64_t now = std::chrono::system_clock::now().time_since_epoch() / std::chrono::seconds(1);
    const std::size_t afs = cwsl_rx_group_af_size(rx, g.id);
    const std::size_t rows = static_cast<std::size_t>(cwsl_rx_num_channels(rx, g.id));
    if (g.audioElems != rows * afs) {  // pinned hand-off buffer, sized for the worst case
        cwsl_host_free(g.audio);
        g.audioElems = rows * afs;
        g.audio = static_cast<std::int16_t*>(cwsl_host_alloc(g.audioElems * sizeof(std::int16_t)));
        if (!g.audio) {
            g.audioElems = 0;
            screenPrinter->err(receiverLog() + std::string("hand-off buffer: ") + cwsl_last_error());
            return;
        }
    }
    // packed hand-off: [decoder][write_index], one contiguous copy over PCIe; the zero tail of each decoder's
    // (period + 5 s) buffer is added below, where the per-decoder vector is built anyway
    std::size_t wi = 0;
    const bool ok = cwsl_rx_end_slot_packed(rx, g.id, g.audio, &wi) == CWSL_OK && cwsl_rx_wait_output(rx) == CWSL_OK;
    const std::uint64_t startTime = g.startEpochTime;
    g.startEpochTime = now;  // stamp of the buffer that starts filling now (Instance.cpp:215)
    if (!ok) {
        screenPrinter->err(receiverLog() + std::string("slot failed: ") + cwsl_last_error());  // failed slot, carry on
    } else if (0 == startTime) {
        screenPrinter->debug(receiverLog() + "Discarding af buffer, start time is zero");  // Instance.cpp:224-227
    } else if (!g.idle) {
        for (std::size_t m = 0; m < g.members.size(); ++m) {
            Instance* inst = g.members[m];
            std::vector<std::int16_t> audioBuf_i16(afs);  // zero-initialised: the tail behind write_index
            std::copy(g.audio + m * wi, g.audio + (m + 1) * wi, audioBuf_i16.begin());
            ItemToDecode toDecode(std::move(audioBuf_i16), inst->getMode(), startTime, inst->getFrequency(),
                                  static_cast<int>(inst->getId()), inst->getCwd(), inst->getTRPeriod());
            inst->getDecoderPool()->push(std::move(toDecode));  // Instance.cpp:244-245
        }
        ++nSlots;
    }
    applyMembership(g);
}

inline void Receiver::readIQ() {
    const std::size_t blockFloats = 2 * iq_len;
    if (!staging) staging = static_cast<float*>(cwsl_host_alloc(kStagingBlocks * blockFloats * sizeof(float)));
    if (!staging) {
        screenPrinter->err(receiverLog() + std::string("IQ staging ring: ") + cwsl_last_error());
        status = ReceiverStatus::STOPPED;
        return;
    }
    std::uint64_t n = 0;  // blocks staged so far
    while (!terminateFlag) {
        // slot edges first, like the Instance does at the top of its loop (Instance.cpp:203-206)
        for (auto& g : groups) {
            if (!g.preds.empty() && g.preds.front()->load()) {
                for (auto& p : g.preds) p->store(false);
                finishSlot(g);
            }
        }
        const std::size_t slot = n % kStagingBlocks;
        // refill a staging buffer only when the copy that last read it has completed (nothing else is waited for)
        if (stagingFence[slot] && cwsl_rx_wait_fence(rx, stagingFence[slot]) != CWSL_OK)
            screenPrinter->err(receiverLog() + std::string("staging fence: ") + cwsl_last_error());
        float* block = staging + slot * blockFloats;
        if (!source->readBlock(block)) {
            screenPrinter->debug(receiverLog() + "IQ producer ended");  // Receiver.hpp:235-237
            break;
        }
        ++n;
        if (cwsl_rx_push_iq(rx, block, 1) != CWSL_OK) {
            // like the reference's "ring buffer full" (Receiver.hpp:222-229): log, drop the block, carry on
            screenPrinter->err(receiverLog() + std::string("push failed, block dropped: ") + cwsl_last_error());
            ++nDroppedBlocks;
            stagingFence[slot] = 0;
            continue;
        }
        if (cwsl_rx_push_fence(rx, &stagingFence[slot]) != CWSL_OK) {
            stagingFence[slot] = 0;
            cwsl_rx_synchronize(rx);  // no fence: fall back to a full wait before this buffer is reused
        }
        ++nBlocks;
    }
    cwsl_rx_synchronize(rx);
    status = ReceiverStatus::STOPPED;
}
