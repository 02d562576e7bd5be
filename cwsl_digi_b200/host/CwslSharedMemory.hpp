// POSIX stand-in for CWSL's Win32 named shared memory.
//
// Same byte layout as source/SharedMemory.h:10-21 and source/SharedMemory.cpp:115-154: page 0 holds
// SM_HDR{int SampleRate, BlockInSamples, L0} followed by DWORD Length, DWORD Current (write offset into
// the circular byte buffer), DWORD LastWrite; the data ring starts at the page boundary. The writer
// advances Current after each block (SharedMemory.cpp:158-202); each reader keeps its own read
// position and may read once `Len` bytes are available (SharedMemory.cpp:207-246). The Win32 named
// event "e"+name (SharedMemory.cpp:134) is replaced by polling Current with a short sleep.
// CwslShmSource adapts a segment to the Receiver's IqSource seam: ONE reader per receiver pushes each
// block to the GPU ring, instead of the reference's one host-ring reader per decoder.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>

#include "IqSource.hpp"

struct SM_HDR {  // source/SharedMemory.h:10-21
    int SampleRate;
    int BlockInSamples;
    int L0;
};

// source/CWSL_Utils.hpp:16-23: "CWSL" + band + "Band" + smNum (POSIX names need a leading '/')
inline std::string createSharedMemName(int bandIndex, int SMNumber) {
    std::string n = "CWSL" + std::to_string(bandIndex) + "Band";
    if (SMNumber != -1) n += std::to_string(SMNumber);
    return n;
}

class CSharedMemory {
public:
    static constexpr std::uint32_t kPage = 4096;
    ~CSharedMemory() { Close(); }

    // writer side (CWSL_Tee in the real system; tests and simulators here)
    bool Create(const std::string& name, std::uint32_t dataLength, const SM_HDR& hdr) {
        Close();
        shmName = "/" + name;
        fd = shm_open(shmName.c_str(), O_CREAT | O_RDWR | O_TRUNC, 0600);
        if (fd < 0) return false;
        length = dataLength + kPage;
        if (ftruncate(fd, length) != 0) return fail();
        if (!map(true)) return fail();
        std::memset(base, 0, kPage);
        *reinterpret_cast<SM_HDR*>(base) = hdr;
        *pLength() = dataLength;
        pCurrent()->store(0);
        owner = write = true;
        init();
        return true;
    }
    bool Open(const std::string& name) {
        Close();
        shmName = "/" + name;
        fd = shm_open(shmName.c_str(), O_RDONLY, 0);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size <= (off_t)kPage) return fail();
        length = static_cast<std::uint32_t>(st.st_size);
        if (!map(false)) return fail();
        init();
        return true;
    }
    void Close() {
        if (base) munmap(base, length);
        if (fd >= 0) close(fd);
        if (owner) shm_unlink(shmName.c_str());
        base = nullptr;
        fd = -1;
        owner = write = false;
    }
    const SM_HDR* GetHeader() const { return reinterpret_cast<const SM_HDR*>(base); }

    bool Write(const std::uint8_t* ptr, std::uint32_t len) {  // SharedMemory.cpp:158-202
        if (!base || !write || len > dataLength) return false;
        const std::uint32_t bLen = dataLength - cur;
        if (len < bLen) {
            std::memcpy(data + cur, ptr, len);
            cur += len;
        } else {
            std::memcpy(data + cur, ptr, bLen);
            std::memcpy(data, ptr + bLen, len - bLen);
            cur = len - bLen;
        }
        pCurrent()->store(cur, std::memory_order_release);
        return true;
    }
    bool Read(std::uint8_t* ptr, std::uint32_t len) {  // SharedMemory.cpp:207-246
        if (!base) return false;
        const std::uint32_t w = pCurrent()->load(std::memory_order_acquire);
        long dLen = static_cast<long>(w) - static_cast<long>(cur);
        if (dLen < 0) dLen += dataLength;
        if (static_cast<long>(len) > dLen) return false;
        const std::uint32_t bLen = dataLength - cur;
        if (len < bLen) {
            std::memcpy(ptr, data + cur, len);
            cur += len;
        } else {
            std::memcpy(ptr, data + cur, bLen);
            std::memcpy(ptr + bLen, data, len - bLen);
            cur = len - bLen;
        }
        return true;
    }
    // replaces WaitForNewData(timeout): poll until `len` bytes can be read
    bool WaitForData(std::uint32_t len, int timeout_ms) {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            const std::uint32_t w = pCurrent()->load(std::memory_order_acquire);
            long dLen = static_cast<long>(w) - static_cast<long>(cur);
            if (dLen < 0) dLen += dataLength;
            if (dLen >= static_cast<long>(len)) return true;
            if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(timeout_ms)) return false;
            std::this_thread::sleep_for(std::chrono::microseconds(200));
        }
    }

private:
    bool map(bool rw) {
        void* p = mmap(nullptr, length, rw ? PROT_READ | PROT_WRITE : PROT_READ, MAP_SHARED, fd, 0);
        if (p == MAP_FAILED) return false;
        base = static_cast<std::uint8_t*>(p);
        return true;
    }
    void init() {
        dataLength = *pLength();
        data = base + kPage;
        cur = pCurrent()->load();
    }
    bool fail() {
        Close();
        return false;
    }
    std::uint32_t* pLength() const { return reinterpret_cast<std::uint32_t*>(base + sizeof(SM_HDR)); }
    std::atomic<std::uint32_t>* pCurrent() const {
        return reinterpret_cast<std::atomic<std::uint32_t>*>(base + sizeof(SM_HDR) + 4);
    }
    std::string shmName;
    int fd = -1;
    std::uint8_t* base = nullptr;
    std::uint8_t* data = nullptr;
    std::uint32_t length = 0, dataLength = 0, cur = 0;
    bool owner = false, write = false;
};

class CwslShmSource : public IqSource {
public:
    explicit CwslShmSource(int timeout_ms = 1000) : timeout(timeout_ms) {}
    bool open(const std::string& smname) override { return SM.Open(smname); }
    std::uint32_t sampleRate() const override { return static_cast<std::uint32_t>(SM.GetHeader()->SampleRate); }
    std::uint32_t blockInSamples() const override { return static_cast<std::uint32_t>(SM.GetHeader()->BlockInSamples); }
    FrequencyHz L0() const override { return static_cast<FrequencyHz>(SM.GetHeader()->L0); }
    bool readBlock(float* dst) override {  // source/Receiver.hpp:233-242
        const std::uint32_t bytes = blockInSamples() * 8u;
        if (!SM.WaitForData(bytes, timeout)) return false;
        return SM.Read(reinterpret_cast<std::uint8_t*>(dst), bytes);
    }

private:
    CSharedMemory SM;
    int timeout;
};
