// Valid-codeword WSPR transmissions for the synthetic IQ source ("valid-codeword signal
// synthesis so a real decoder, when available, has something to decode").
//
// The reference only ever RECEIVES: every 120 s WSPR slot goes to the external wsprd as a WAV file
// (source/DecoderPool.hpp:1007-1026; period table source/CWSL_DIGI.hpp:64-113). To give that hand-off something a
// decoder accepts, this header builds the transmit side of the protocol from its published constants: type-1 message
// packing (callsign 28 bits, locator + power 22 bits), the K = 32, r = 1/2 convolutional code (polynomials
// 0xF2D05351 / 0xE4613C47, 31 zero tail bits), the bit-reversal interleaver, the 162-bit sync vector, and the air
// interface (4-FSK, 12000/8192 = 1.4648 Hz tone spacing and baud, continuous phase, start 1 s into the even minute).
// Validated, not recollected: the encoder reproduces all 162 published channel symbols of the protocol's standard
// example "K1ABC FN42 37" (host_tests.cpp; the same known-answer test pins tests/wspr_codec.py, whose blind decoder
// then recovers the messages from the audio the GPU front-end produces).
// FT8 / FT4 are NOT here: their LDPC(174,91) generator matrix cannot be derived or validated offline.
#pragma once

#include <array>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

namespace wspr {

constexpr int kSymbols = 162;
constexpr double kToneHz = 12000.0 / 8192.0;   // tone spacing = symbol rate
constexpr double kSymbolS = 8192.0 / 12000.0;  // 0.6827 s

// sync vector, packed LSB-first into bytes in transmission order
inline bool syncBit(int i) {
    static const std::uint8_t v[21] = {0x03, 0x71, 0xA4, 0x07, 0xA4, 0x40, 0xB3, 0x58, 0x58, 0x95, 0x34, 0x56,
                                       0x04, 0xC9, 0xCD, 0xE2, 0xA0, 0x0C, 0x58, 0x63, 0x00};
    return (v[i >> 3] >> (i & 7)) & 1;
}

inline int popParity(std::uint32_t x) {
    x ^= x >> 16;
    x ^= x >> 8;
    x ^= x >> 4;
    return (0x6996u >> (x & 15u)) & 1;
}

// The 162 channel symbols (0..3) of a type-1 message. false (with *why set) when the callsign does not fit the
// 6-character template (digit in third place), the locator is not a 4-character Maidenhead square, or the power
// is not one of the 19 WSPR levels (0..60 dBm ending in 0, 3 or 7).
inline bool encode(const std::string& call_in, const std::string& grid_in, int dbm, std::array<std::uint8_t, kSymbols>& out,
                   std::string* why = nullptr) {
    auto bad = [&](const char* m) {
        if (why) *why = m;
        return false;
    };
    std::string call, grid;
    for (char ch : call_in)
        if (ch != ' ') call += static_cast<char>(std::toupper(static_cast<unsigned char>(ch)));
    for (char ch : grid_in) grid += static_cast<char>(std::toupper(static_cast<unsigned char>(ch)));
    auto isdig = [](char ch) { return ch >= '0' && ch <= '9'; };
    auto isup = [](char ch) { return ch >= 'A' && ch <= 'Z'; };
    if (call.size() >= 2 && isdig(call[1]) && !(call.size() >= 3 && isdig(call[2]))) call = " " + call;
    while (call.size() < 6) call += ' ';
    if (call.size() != 6 || !isdig(call[2])) return bad("callsign does not fit a type-1 message");
    auto code = [&](char ch) -> int { return isdig(ch) ? ch - '0' : ch == ' ' ? 36 : isup(ch) ? ch - 'A' + 10 : -1; };
    int c[6];
    for (int i = 0; i < 6; ++i) {
        c[i] = code(call[i]);
        if (c[i] < 0 || (i >= 3 && c[i] < 10) || (i == 1 && c[i] == 36)) return bad("bad callsign character");
    }
    std::uint32_t n = static_cast<std::uint32_t>(c[0]);
    n = n * 36 + c[1];
    n = n * 10 + c[2];
    for (int i = 3; i < 6; ++i) n = n * 27 + (c[i] - 10);  // letters 0..25, space 26
    if (grid.size() != 4 || grid[0] < 'A' || grid[0] > 'R' || grid[1] < 'A' || grid[1] > 'R' || !isdig(grid[2]) || !isdig(grid[3]))
        return bad("bad locator");
    const std::uint32_t m1 = (179 - 10 * (grid[0] - 'A') - (grid[2] - '0')) * 180 + 10 * (grid[1] - 'A') + (grid[3] - '0');
    if (dbm < 0 || dbm > 60 || !(dbm % 10 == 0 || dbm % 10 == 3 || dbm % 10 == 7)) return bad("not a WSPR power level");
    const std::uint32_t m = m1 * 128 + static_cast<std::uint32_t>(dbm) + 64;
    // 50 message bits (N then M, MSB first) + 31 zero tail bits through the convolutional encoder
    std::uint8_t coded[kSymbols];
    std::uint32_t reg = 0;
    for (int i = 0; i < 81; ++i) {
        const int bit = i < 28 ? (n >> (27 - i)) & 1 : i < 50 ? (m >> (49 - i)) & 1 : 0;
        reg = (reg << 1) | static_cast<std::uint32_t>(bit);
        coded[2 * i] = static_cast<std::uint8_t>(popParity(reg & 0xF2D05351u));
        coded[2 * i + 1] = static_cast<std::uint8_t>(popParity(reg & 0xE4613C47u));
    }
    // interleave: coded bit p goes to the p-th position j = bitreverse8(i) < 162, i ascending
    int p = 0;
    for (int i = 0; i < 256; ++i) {
        int j = 0;
        for (int b = 0; b < 8; ++b)
            if (i & (1 << b)) j |= 0x80 >> b;
        if (j < kSymbols) out[j] = static_cast<std::uint8_t>((syncBit(j) ? 1 : 0) + 2 * coded[p++]);
    }
    return true;
}

}  // namespace wspr

// One M-FSK transmission for SyntheticIqSource: continuous phase, `symbols[k]` selects tone rf_hz + symbols[k] * tone_hz
// during [t0_s + k symbol_s, t0_s + (k+1) symbol_s) of every `period_s` (0: once, counted from the stream's start).
struct FskBurst {
    double rf_hz = 0;      // absolute RF frequency of tone 0
    double tone_hz = wspr::kToneHz;
    double symbol_s = wspr::kSymbolS;
    double t0_s = 1.0;
    double period_s = 120.0;
    double amplitude = 0;
    std::vector<std::uint8_t> symbols;
    double phase = 0;      // running phase, cycles (state of the source)
};

// A WSPR transmission whose four tones are centred on `centre_rf_hz` (dial + 1500 Hz is the middle of wsprd's band).
inline bool makeWsprBurst(const std::string& call, const std::string& grid, int dbm, double centre_rf_hz, double amplitude,
                          FskBurst& b, double t0_s = 1.0, std::string* why = nullptr) {
    std::array<std::uint8_t, wspr::kSymbols> s{};
    if (!wspr::encode(call, grid, dbm, s, why)) return false;
    b = FskBurst{};
    b.rf_hz = centre_rf_hz - 1.5 * wspr::kToneHz;
    b.t0_s = t0_s;
    b.amplitude = amplitude;
    b.symbols.assign(s.begin(), s.end());
    return true;
}

// Amplitude of a complex exponential that is snr_db (in 2500 Hz, the bandwidth WSJT-X quotes SNRs in) above complex
// white noise of standard deviation sigma per component at sample rate fs.
inline double amplitudeForSnr(double snr_db_2500, double sigma, double fs) {
    return std::sqrt(2.0 * sigma * sigma * 2500.0 / fs * std::pow(10.0, snr_db_2500 / 10.0));
}
