// Host-side mirror of the reference's shared types for the receive front-end.
// Same names and meaning as source/CWSL_DIGI_Types.hpp:33-145 and source/CWSL_DIGI.hpp:44-113 so
// that code written against the reference's Receiver / Decoder / DecoderPool keeps compiling.
#pragma once

#include <atomic>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

using FrequencyHz = std::uint32_t;  // source/CWSL_DIGI_Types.hpp:33

constexpr float Q65_30_PERIOD = 30.0f;  // source/CWSL_DIGI.hpp:44-49
constexpr float FT8_PERIOD = 15.0f;
constexpr float FT4_PERIOD = 7.5f;
constexpr float WSPR_PERIOD = 120.0f;
constexpr float JT65_PERIOD = 60.0f;
constexpr float JS8_PERIOD = 15.0f;

constexpr size_t Wave_SR = 12000;  // source/CWSL_DIGI.hpp:51-53
constexpr size_t SSB_BW = 6000;
constexpr bool USB = 1;

// source/CWSL_DIGI.hpp:64-113
static inline float getRXPeriod(const std::string& mode) {
    static const std::map<std::string, float> periods = {
        {"FT8", FT8_PERIOD},     {"JS8", JS8_PERIOD},      {"FT4", FT4_PERIOD},      {"WSPR", WSPR_PERIOD},
        {"Q65-30", Q65_30_PERIOD}, {"JT65", JT65_PERIOD},  {"FST4-60", 60.0f},       {"FST4-120", 120.0f},
        {"FST4-300", 300.0f},    {"FST4-900", 900.0f},     {"FST4-1800", 1800.0f},   {"FST4W-120", 120.0f},
        {"FST4W-300", 300.0f},   {"FST4W-900", 900.0f},    {"FST4W-1800", 1800.0f}};
    auto it = periods.find(mode);
    if (it == periods.end()) throw std::runtime_error("Unhandled mode: " + mode);
    return it->second;
}

// The modes a decoder= line may name (source/CWSL_DIGI.cpp:744-803; Q65-30 is accepted there too)
static inline bool isKnownMode(const std::string& mode) {
    try {
        (void)getRXPeriod(mode);
        return true;
    } catch (const std::runtime_error&) {
        return false;
    }
}

enum class InstanceStatus : int { NOT_INITIALIZED, RUNNING, STOPPED, FINISHED };  // CWSL_DIGI_Types.hpp:57-62
enum class ReceiverStatus : int { NOT_INITIALIZED, RUNNING, STOPPED, FINISHED };  // source/Receiver.hpp:45-50

// Slot-edge flag, set by a clock thread, consumed by the front-end at the next IQ block
// (source/CWSL_DIGI_Types.hpp:65-78).
class SyncPredicate {
public:
    SyncPredicate() { pred.store(false); }
    bool load() { return pred.load(); }
    void store(bool a) { pred.store(a); }

private:
    std::atomic_bool pred;
};

// One predicate per decoder, kept in per-period lists; a clock thread sets every predicate of its
// period at the slot edge (source/CWSL_DIGI_Types.hpp:80-145: ft8Preds holds FT8 and JS8, s60sPreds
// JT65 and FST4-60, s120sPreds WSPR/FST4-120/FST4W-120, ...; source/CWSL_DIGI.cpp:234-262).
// The decoders of one receiver that share a period form one "slot group" on the GPU.
class SyncPredicates {
public:
    std::shared_ptr<SyncPredicate> createPredicate(const std::string& mode) {
        const float period = getRXPeriod(mode);  // throws "Unhandled mode" like the reference
        std::shared_ptr<SyncPredicate> pred = std::make_shared<SyncPredicate>();
        std::lock_guard<std::mutex> lk(mu);
        byPeriod[period].push_back(pred);
        return pred;
    }
    // what one waitForTime* thread does at its slot edge: preds[k]->store(true) for all k
    void fire(float period_s) {
        std::lock_guard<std::mutex> lk(mu);
        auto it = byPeriod.find(period_s);
        if (it == byPeriod.end()) return;
        for (auto& p : it->second) p->store(true);
    }
    std::vector<float> periods() {
        std::lock_guard<std::mutex> lk(mu);
        std::vector<float> v;
        for (auto& kv : byPeriod) v.push_back(kv.first);
        return v;
    }

private:
    std::mutex mu;
    std::map<float, std::vector<std::shared_ptr<SyncPredicate>>> byPeriod;
};

// Minimal levelled logger with the reference's method names (source/ScreenPrinter.hpp:37-45);
// the asynchronous print thread and log file are out of scope.
enum class LOG_LEVEL : int { DEBUG = 0, INFO = 1, WARN = 2, ERR = 3 };
class ScreenPrinter {
public:
    explicit ScreenPrinter(LOG_LEVEL lvl = LOG_LEVEL::INFO) : level(lvl) {}
    void print(const std::string& s, LOG_LEVEL l = LOG_LEVEL::INFO) {
        if (static_cast<int>(l) < static_cast<int>(level)) return;
        std::lock_guard<std::mutex> lk(mu);
        std::fprintf(l == LOG_LEVEL::ERR ? stderr : stdout, "%s\n", s.c_str());
    }
    void print(const std::string& s, const std::exception& e) { print(s + ": " + e.what(), LOG_LEVEL::ERR); }
    void debug(const std::string& s) { print(s, LOG_LEVEL::DEBUG); }
    void info(const std::string& s) { print(s, LOG_LEVEL::INFO); }
    void warning(const std::string& s) { print(s, LOG_LEVEL::WARN); }
    void err(const std::string& s) { print(s, LOG_LEVEL::ERR); }

private:
    LOG_LEVEL level;
    std::mutex mu;
};
