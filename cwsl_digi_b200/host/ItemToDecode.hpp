// The hand-off record between a decoder Instance and the DecoderPool (source/DecoderPool.hpp:174-210).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "CWSL_DIGI_Types.hpp"

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
