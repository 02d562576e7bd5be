// config.ini `decoder=` semantics and the hot-path knobs of the [wsjtx] section.
//
// Grammar and checks follow source/CWSL_DIGI.cpp:731-842 exactly:
//   decoder = <freq_hz> <mode> [<sharedmem> [<freqcal> [<callsign>]]]
// split on single spaces into 2..5 tokens; mode must be one of the known modes (:744-803); the
// callsign token is only accepted for WSPR (:827-833); calibrated frequency =
// uint32(freq / (freqcalibration * decoder_freqcal)) (:834). wsjtx.ftaudioscalefactor and
// wsjtx.wspraudioscalefactor must be in (0, 1] with defaults 0.90 / 0.20 (:100-101, :952-978).
// The rest of the option table (reporting, logging, process management) is out of scope.
#pragma once

#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "CWSL_DIGI_Types.hpp"
#include "Decoder.hpp"

// source/StringUtils.hpp:51-71 (noEmpty = false): consecutive delimiters yield empty tokens
static inline std::vector<std::string> splitStringByDelim(const std::string& input, const char delim) {
    std::istringstream iss(input);
    std::vector<std::string> v;
    std::string s;
    while (std::getline(iss, s, delim)) v.push_back(s);
    return v;
}

struct FrontEndConfig {
    DecoderVec decoders;
    double freqCalGlobal = 1.0;          // radio.freqcalibration
    int SMNumber = -1;                   // radio.sharedmem (-1 = search, source/CWSL_Utils.hpp:27-53)
    float ftAudioScaleFactor = 0.90f;    // wsjtx.ftaudioscalefactor
    float wsprAudioScaleFactor = 0.20f;  // wsjtx.wspraudioscalefactor
    std::string operatorCallsign;        // operator.callsign
    // Extension (not in the reference's config.ini; the reference parses with allow_unregistered = true,
    // source/CWSL_DIGI.cpp:607, so it skips them): [gpu] arithmetic=exact|fast|stft, device=N. Values are the CWSL_MODE_* codes of include/cwsl_b200.h.
    int kernelMode = 1;                  // gpu.arithmetic, default fast
    int cudaDevice = 0;                  // gpu.device
};

// Parse one decoder= value. Throws std::invalid_argument with the reference's message prefix.
inline Decoder parseDecoderLine(const std::string& rawLine, double freqCalGlobal, int defaultSmNum,
                                const std::string& operatorCallsign) {
    const auto tok = splitStringByDelim(rawLine, ' ');
    if (tok.size() < 2 || tok.size() > 5) throw std::invalid_argument("Error parsing decoder line: " + rawLine);
    const uint32_t freq = static_cast<uint32_t>(std::stoi(tok[0]));
    const std::string& mode = tok[1];
    if (!isKnownMode(mode))
        throw std::invalid_argument("Error parsing decoder line, unknown mode: " + mode + "Full Line: " + rawLine);
    int smnum = defaultSmNum;
    if (tok.size() >= 3) smnum = std::stoi(tok[2]);
    double decoder_freqcal = 1.0;
    if (tok.size() >= 4) decoder_freqcal = std::stod(tok[3]);
    std::string callsign = operatorCallsign;
    if (tok.size() >= 5) {
        if (mode != "WSPR")
            throw std::invalid_argument("Callsigns are only supported per-decoder for WSPR decoders");
        callsign = tok[4];
    }
    const FrequencyHz calibrated = static_cast<FrequencyHz>(freq / (freqCalGlobal * decoder_freqcal));
    return Decoder(freq, calibrated, mode, smnum, decoder_freqcal, callsign);
}

// Minimal INI reader for the keys above ([section] headers, key=value, '#' starts a comment anywhere in a line), the
// format boost::program_options::parse_config_file accepts (source/CWSL_DIGI.cpp:607-611). ';' is NOT a comment
// character there, so it is not one here either.
inline FrontEndConfig loadFrontEndConfig(std::istream& in) {
    FrontEndConfig cfg;
    std::vector<std::string> decoderLines;
    std::string line, section;
    auto trim = [](std::string s) {
        const char* ws = " \t\r\n";
        const auto b = s.find_first_not_of(ws);
        if (b == std::string::npos) return std::string();
        return s.substr(b, s.find_last_not_of(ws) - b + 1);
    };
    while (std::getline(in, line)) {
        const auto hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        line = trim(line);
        if (line.empty()) continue;
        if (line.front() == '[' && line.back() == ']') {
            section = line.substr(1, line.size() - 2);
            continue;
        }
        const auto eq = line.find('=');
        if (eq == std::string::npos) continue;
        const std::string key = section + "." + trim(line.substr(0, eq));
        const std::string val = trim(line.substr(eq + 1));
        if (key == "decoders.decoder") decoderLines.push_back(val);
        else if (key == "radio.freqcalibration") cfg.freqCalGlobal = std::stod(val);
        else if (key == "radio.sharedmem") cfg.SMNumber = std::stoi(val);
        else if (key == "operator.callsign") cfg.operatorCallsign = val;
        else if (key == "wsjtx.ftaudioscalefactor") cfg.ftAudioScaleFactor = std::stof(val);
        else if (key == "wsjtx.wspraudioscalefactor") cfg.wsprAudioScaleFactor = std::stof(val);
        else if (key == "gpu.device") cfg.cudaDevice = std::stoi(val);
        else if (key == "gpu.arithmetic") {
            if (val == "exact") cfg.kernelMode = 0;
            else if (val == "fast") cfg.kernelMode = 1;
            else if (val == "stft") cfg.kernelMode = 2;
            else throw std::invalid_argument("gpu.arithmetic must be exact, fast or stft");
        }
    }
    if (cfg.ftAudioScaleFactor > 1.0f || cfg.ftAudioScaleFactor <= 0.0f)  // source/CWSL_DIGI.cpp:952-964
        throw std::invalid_argument("ftaudioscalefactor must be > 0 and <= 1");
    if (cfg.wsprAudioScaleFactor > 1.0f || cfg.wsprAudioScaleFactor <= 0.0f)  // :966-978
        throw std::invalid_argument("wspraudioscalefactor must be > 0 and <= 1");
    if (decoderLines.empty())
        throw std::invalid_argument("decoders.decoder input is required but was not specified!");  // :838
    for (const auto& l : decoderLines)
        cfg.decoders.push_back(parseDecoderLine(l, cfg.freqCalGlobal, cfg.SMNumber, cfg.operatorCallsign));
    return cfg;
}

inline FrontEndConfig loadFrontEndConfigFile(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::invalid_argument("cannot open config file: " + path);
    return loadFrontEndConfig(f);
}
