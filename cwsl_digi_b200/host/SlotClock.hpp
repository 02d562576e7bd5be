// Wall-clock slot edges: what the reference's eight waitForTime* threads do
// (source/CWSL_DIGI.cpp:174-451, started at :1134-1175). One polling thread per period in use sets every
// SyncPredicate of that period when UTC crosses the period's edge:
//   7.5 s  FT4      sec in {0,15,30,45}, and sec in {7,22,37,52} once ms >= 300      (:402-451)
//   15 s   FT8/JS8  sec in {0,15,30,45}                                              (:234-262)
//   30 s   Q65-30   sec in {0,30}                                                    (:174-202)
//   60 s            sec == 0                                                         (:204-232)
//   120 s           even minute, sec == 0                                            (:367-399)
//   300/900/1800 s  minute % 5/15/30 == 0, sec == 0                                  (:264-364)
// An edge fires once: the thread latches the (minute, second) it fired on, like `goSec`.
// The clock is injectable so the rule table is unit-testable without waiting.
#pragma once

#include <atomic>
#include <chrono>
#include <cmath>
#include <ctime>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include "CWSL_DIGI_Types.hpp"

constexpr int MIN_SLEEP_MS = 25;  // source/CWSL_DIGI.hpp:58

struct UtcTime {
    int minute = 0, second = 0, millis = 0;
    static UtcTime fromEpochMs(std::uint64_t ms) {
        UtcTime t;
        t.millis = static_cast<int>(ms % 1000);
        const std::uint64_t s = ms / 1000;
        t.second = static_cast<int>(s % 60);
        t.minute = static_cast<int>((s / 60) % 60);
        return t;
    }
};

// true while UTC sits on a slot edge of `period_s` (the `go` expression of the matching thread)
inline bool slotEdgeNow(float period_s, const UtcTime& t) {
    const int s = t.second, m = t.minute;
    const bool q = (s == 0 || s == 15 || s == 30 || s == 45);
    if (std::fabs(period_s - 7.5f) < 1e-3f) return q || ((s == 7 || s == 22 || s == 37 || s == 52) && t.millis >= 300);
    if (std::fabs(period_s - 15.0f) < 1e-3f) return q;
    if (std::fabs(period_s - 30.0f) < 1e-3f) return s == 0 || s == 30;
    if (std::fabs(period_s - 60.0f) < 1e-3f) return s == 0;
    if (std::fabs(period_s - 120.0f) < 1e-3f) return (m & 1) == 0 && s == 0;
    if (std::fabs(period_s - 300.0f) < 1e-3f) return m % 5 == 0 && s == 0;
    if (std::fabs(period_s - 900.0f) < 1e-3f) return m % 15 == 0 && s == 0;
    if (std::fabs(period_s - 1800.0f) < 1e-3f) return m % 30 == 0 && s == 0;
    return false;
}

class SlotClocks {
public:
    using NowFn = std::function<std::uint64_t()>;  // epoch milliseconds (UTC)
    explicit SlotClocks(std::shared_ptr<SyncPredicates> predsIn, NowFn nowIn = nullptr)
        : preds(std::move(predsIn)), now(nowIn ? std::move(nowIn) : NowFn(&SlotClocks::systemNowMs)) {}
    ~SlotClocks() { stop(); }

    static std::uint64_t systemNowMs() {
        return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch())
            .count();
    }

    // one thread per period that has at least one predicate (source/CWSL_DIGI.cpp:1134-1175)
    void start() {
        terminateFlag = false;
        for (float p : preds->periods()) threads.emplace_back(&SlotClocks::run, this, p);
    }
    void stop() {
        terminateFlag = true;
        for (auto& t : threads)
            if (t.joinable()) t.join();
        threads.clear();
    }

    // One poll of the rule for `period`: fires at most once per (minute, second). Returns true if fired.
    // `latch` carries the thread's goSec-style state between polls.
    bool poll(float period, int& latch) {
        const UtcTime t = UtcTime::fromEpochMs(now());
        const int key = t.minute * 60 + t.second;
        if (!slotEdgeNow(period, t)) {
            if (latch != key) latch = -1;
            return false;
        }
        if (latch == key) return false;
        latch = key;
        preds->fire(period);
        return true;
    }

private:
    void run(float period) {
        int latch = -1;
        while (!terminateFlag) {
            poll(period, latch);
            std::this_thread::sleep_for(std::chrono::milliseconds(MIN_SLEEP_MS));
        }
    }
    std::shared_ptr<SyncPredicates> preds;
    NowFn now;
    std::vector<std::thread> threads;
    std::atomic_bool terminateFlag{false};
};
