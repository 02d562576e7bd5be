// Out-of-line members of Decoder that need the complete Instance type (source/Decoder.cpp:23-86).
#pragma once
#include "Decoder.hpp"
#include "Instance.hpp"

inline Decoder::Decoder(FrequencyHz freq_In, FrequencyHz freqCalibrated_In, std::string mode_In, int smNum_In,
                        double freqCalFactor_In, std::string reporterCallsign_In)
    : freq(freq_In), freqCalibrated(freqCalibrated_In), mode(std::move(mode_In)), smNum(smNum_In),
      freqCalFactor(freqCalFactor_In), reporterCallsign(std::move(reporterCallsign_In)), instance(nullptr) {}
inline Decoder::Decoder(Decoder&&) noexcept = default;
inline Decoder& Decoder::operator=(Decoder&&) noexcept = default;
inline Decoder::~Decoder() = default;
inline void Decoder::setInstance(std::unique_ptr<Instance> inst) { instance = std::move(inst); }
inline InstanceStatus Decoder::getStatus() {
    return instance ? instance->getStatus() : InstanceStatus::NOT_INITIALIZED;
}
inline void Decoder::terminate() {
    if (instance) instance->terminate();
}
