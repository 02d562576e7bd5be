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
    void terminate() { status = InstanceStatus::FINISHED; }
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
inline bool Receiver::addInstance(Instance* inst) {
    if (!rx || status == ReceiverStatus::RUNNING) return false;
    SlotGroup* g = nullptr;
    for (auto& sg : groups)
        if (sg.period == inst->getTRPeriod()) g = &sg;
    if (!g) {
        SlotGroup sg;
        sg.period = inst->getTRPeriod();
        sg.id = cwsl_rx_add_group(rx, sg.period);
        if (sg.id < 0) return false;
        groups.push_back(std::move(sg));
        g = &groups.back();
    }
    const int ch = cwsl_rx_add_channel(rx, g->id, inst->demodFreq(), USB, inst->audioScale());
    if (ch < 0) return false;  // e.g. "Signal outside of band", source/SSBD.hpp:100-103
    inst->group = g->id;
    inst->channel = ch;
    g->members.push_back(inst);
    g->preds.push_back(inst->getPredicate());
    instances.push_back(inst);
    return true;
}

inline void Receiver::finishSlot(SlotGroup& g) {
    // source/Instance.cpp:203-253 for every decoder of the group
    const std::uint64_t now = std::chrono::system_clock::now().time_since_epoch() / std::chrono::seconds(1);
    const std::size_t afs = cwsl_rx_group_af_size(rx, g.id);
    if (g.audioElems != g.members.size() * afs) {  // pinned + managed: only the demodulated columns cross PCIe
        cwsl_host_free(g.audio);
        g.audioElems = g.members.size() * afs;
        g.audio = static_cast<std::int16_t*>(cwsl_host_alloc(g.audioElems * sizeof(std::int16_t)));
        if (!g.audio) {
            g.audioElems = 0;
            screenPrinter->err(receiverLog() + std::string("hand-off buffer: ") + cwsl_last_error());
            return;
        }
    }
    std::size_t wi = 0;
    if (cwsl_rx_end_slot(rx, g.id, g.audio, &wi) != CWSL_OK || cwsl_rx_wait_output(rx) != CWSL_OK) {
        screenPrinter->err(receiverLog() + std::string("slot failed: ") + cwsl_last_error());  // failed slot, carry on
        g.startEpochTime = now;
        return;
    }
    const std::uint64_t startTime = g.startEpochTime;
    g.startEpochTime = now;  // stamp of the buffer that starts filling now (Instance.cpp:215)
    if (0 == startTime) {
        screenPrinter->debug(receiverLog() + "Discarding af buffer, start time is zero");  // Instance.cpp:224-227
        return;
    }
    for (std::size_t m = 0; m < g.members.size(); ++m) {
        Instance* inst = g.members[m];
        std::vector<std::int16_t> audioBuf_i16(g.audio + m * afs, g.audio + (m + 1) * afs);
        ItemToDecode toDecode(std::move(audioBuf_i16), inst->getMode(), startTime, inst->getFrequency(),
                              static_cast<int>(inst->getId()), inst->getCwd(), inst->getTRPeriod());
        inst->getDecoderPool()->push(std::move(toDecode));  // Instance.cpp:244-245
    }
    ++nSlots;
}

inline void Receiver::readIQ() {
    std::vector<float> block(2 * iq_len);
    while (!terminateFlag) {
        // slot edges first, like the Instance does at the top of its loop (Instance.cpp:203-206)
        for (auto& g : groups) {
            if (g.preds.front()->load()) {
                for (auto& p : g.preds) p->store(false);
                finishSlot(g);
            }
        }
        if (!source->readBlock(block.data())) {
            screenPrinter->debug(receiverLog() + "IQ producer ended");  // Receiver.hpp:235-237
            break;
        }
        if (cwsl_rx_push_iq(rx, block.data(), 1) != CWSL_OK) {
            screenPrinter->err(receiverLog() + std::string("push failed: ") + cwsl_last_error());
            break;
        }
        // pageable source buffer is reused for the next block: wait for the copy
        cwsl_rx_synchronize(rx);
        ++nBlocks;
    }
    status = ReceiverStatus::STOPPED;
}
