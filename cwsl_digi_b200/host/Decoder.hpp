// Value object for one `decoder=` line; interface of source/Decoder.hpp:31-69.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "CWSL_DIGI_Types.hpp"

class Instance;

class Decoder {
public:
    Decoder(FrequencyHz freq_In, FrequencyHz freqCalibrated_In, std::string mode_In, int smNum_In,
            double freqCalFactor_In, std::string reporterCallsign_In);
    Decoder(Decoder&&) noexcept;
    Decoder& operator=(Decoder&&) noexcept;
    ~Decoder();

    float getTRPeriod() { return getRXPeriod(mode); }
    void setInstance(std::unique_ptr<Instance> inst);
    InstanceStatus getStatus();
    void terminate();
    int getsmNum() const { return smNum; }
    FrequencyHz getFreq() const { return freq; }
    FrequencyHz getFreqCalibrated() const { return freqCalibrated; }
    std::string getMode() const { return mode; }
    double getFreqCalFactor() const { return freqCalFactor; }
    std::string getReporterCallsign() const { return reporterCallsign; }
    std::unique_ptr<Instance>& getInstance() { return instance; }

private:
    FrequencyHz freq;
    FrequencyHz freqCalibrated;
    std::string mode;
    int smNum;
    double freqCalFactor;
    std::string reporterCallsign;
    std::unique_ptr<Instance> instance;
};

using DecoderVec = std::vector<Decoder>;
