// WAV hand-off file for jt9/wsprd.
// Byte-for-byte the file the reference writes: a packed 46-byte header -- "RIFF", u32 FileLen =
// 46 + DataLen - 8, "WAVE", "fmt ", u32 FmtLen = sizeof(WAVEFORMATEX) = 18, {PCM=1, 1 channel,
// 12000 Hz, 24000 B/s, block align 2, 16 bit, cbSize 0}, "data", u32 DataLen -- followed by the
// int16 samples of the whole (period+5 s) buffer (source/WaveFile.hpp:19-35, :87-134;
// source/DecoderPool.hpp:915-964).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#pragma pack(push, 1)
struct WavHdr {
    char _RIFF[4];
    std::uint32_t FileLen;
    char _WAVE[4];
    char _fmt[4];
    std::uint32_t FmtLen;
    struct {  // WAVEFORMATEX (18 bytes)
        std::uint16_t wFormatTag;
        std::uint16_t nChannels;
        std::uint32_t nSamplesPerSec;
        std::uint32_t nAvgBytesPerSec;
        std::uint16_t nBlockAlign;
        std::uint16_t wBitsPerSample;
        std::uint16_t cbSize;
    } Format;
    char _data[4];
    std::uint32_t DataLen;
};
#pragma pack(pop)
static_assert(sizeof(WavHdr) == 46, "the reference's WAV header is 46 bytes");

inline WavHdr makeWavHdr(std::size_t n_samples) {
    const std::size_t DataLen = n_samples * sizeof(std::int16_t);
    WavHdr h;
    std::memcpy(h._RIFF, "RIFF", 4);
    h.FileLen = static_cast<std::uint32_t>((sizeof(h) + DataLen) - 8);
    std::memcpy(h._WAVE, "WAVE", 4);
    std::memcpy(h._fmt, "fmt ", 4);
    h.FmtLen = 18;
    h.Format.wFormatTag = 1;  // WAVE_FORMAT_PCM
    h.Format.nChannels = 1;
    h.Format.nSamplesPerSec = 12000;
    h.Format.nBlockAlign = 2;
    h.Format.nAvgBytesPerSec = h.Format.nSamplesPerSec * h.Format.nBlockAlign;
    h.Format.wBitsPerSample = 16;
    h.Format.cbSize = 0;
    std::memcpy(h._data, "data", 4);
    h.DataLen = static_cast<std::uint32_t>(DataLen);
    return h;
}

// false on I/O error (the reference logs and carries on, source/DecoderPool.hpp:949-956)
inline bool waveWrite(const std::vector<std::int16_t>& audioBuffer, const std::string& fileName) {
    const WavHdr h = makeWavHdr(audioBuffer.size());
    FILE* f = std::fopen(fileName.c_str(), "wb");
    if (!f) return false;
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1;
    ok = ok && (audioBuffer.empty() ||
                std::fwrite(audioBuffer.data(), sizeof(std::int16_t), audioBuffer.size(), f) == audioBuffer.size());
    return (std::fclose(f) == 0) && ok;
}
