// jt9 shared-memory hand-off: the segment layout and the per-mode parameter fill of
// DecoderPool::decodeUsingShMem (source/DecoderPool.hpp:44-108 layout, :421-593 fill, :575-577 handshake).
// `dec_data_t` is shared with WSJT-X's Fortran (lib/jt9com.f90) and must stay in sync with it.
// The reference creates the segment with Qt's QSharedMemory (a CreateFileMapping keyed by a Qt-mangled name on
// Windows, source/DecoderPool.hpp:422-437); Jt9ShmSegment is its POSIX stand-in (shm_open + mmap under "/<key>"),
// with the same life cycle: create -> fill -> decoder attaches by key -> handshake -> detach. Spawning jt9 itself
// stays out of scope (Win32 CreateProcessA, binary not available). fillDecData() works on any block of
// sizeof(dec_data_t) bytes (the mapped segment, or pinned memory the GPU result was copied into).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>

#include "ItemToDecode.hpp"

#define NSMAX 6827
#define NTMAX (30 * 60)
#define RX_SAMPLE_RATE 12000

typedef struct dec_data {
    int ipc[3];                      // [0] nzhsym, [1] istart, [2] idone handshake words
    float ss[184 * NSMAX];           // symbol spectra (jt9 fills/reads; zeroed here)
    float savg[NSMAX];
    float sred[5760];
    short int d2[NTMAX * RX_SAMPLE_RATE];  // 12 kHz int16 audio: where the front-end's output goes
    struct {                         // field order and types are the ABI of lib/jt9com.f90 -- do not reorder
        int nutc;                    // HHMM
        bool ndiskdat;               // data came from a .wav file
        int ntrperiod;               // T/R period, s
        int nQSOProgress;
        int nfqso;                   // QSO audio frequency
        int nftx;
        bool newdat;                 // new data: run the long FFT
        int npts8;
        int nfa;                     // lowest decode frequency, Hz
        int nfSplit;
        int nfb;                     // highest decode frequency, Hz
        int ntol;                    // search range around nfqso, Hz
        int kin;
        int nzhsym;                  // half-symbol count that triggers the decode
        int nsubmode;
        bool nagain;
        int ndepth;                  // decode depth 1..3
        bool lft8apon;
        bool lapcqonly;
        bool ljt65apon;
        int napwid;
        int ntxmode;
        int nmode;                   // 8 FT8, 5 FT4, 65 JT65, 66 Q65, 240 FST4, 241 FST4W
        int minw;
        bool nclearave;
        int minSync;
        float emedelay;
        float dttol;
        int nlist;
        int listutc[10];
        int n2pass;
        int nranera;
        int naggressive;
        bool nrobust;
        int nexp_decode;
        char datetime[20];
        char mycall[12];
        char mygrid[6];
        char hiscall[12];
        char hisgrid[6];
    } params;
} dec_data_t;

static_assert(offsetof(dec_data_t, ss) == 12, "ipc[3] precedes ss");
static_assert(offsetof(dec_data_t, d2) == 12 + 4 * (184 * NSMAX + NSMAX + 5760), "d2 offset (jt9com.f90)");
static_assert(offsetof(dec_data_t, params) == offsetof(dec_data_t, d2) + 2 * NTMAX * RX_SAMPLE_RATE, "params follow d2");

// Fills the whole block exactly as the reference does. false for a mode jt9 does not take over shared memory
// (WSPR/JS8 always go through WAV files, source/DecoderPool.hpp:379-392) or an unknown mode.
inline bool fillDecData(dec_data_t* dec_data, const ItemToDecode& item, int highestDecodeFreq, int decodedepth) {
    std::memset(dec_data, 0, sizeof(dec_data_t));
    auto& p = dec_data->params;
    p.nfa = 0;
    p.nfb = highestDecodeFreq;
    p.ndepth = decodedepth;
    p.nutc = 0;
    p.newdat = 1;
    p.nagain = 0;
    p.emedelay = 0;
    p.nrobust = 0;
    p.ndiskdat = 0;
    p.minw = 0;
    p.minSync = 0;
    p.dttol = 4;
    const std::string& m = item.mode;
    auto fst4 = [&](int nfa, int nzhsym, int period) {
        p.ndepth = 1; p.nfa = nfa; p.nfb = 1100; p.nzhsym = nzhsym; p.nmode = 240; p.ntol = 100; p.ntrperiod = period;
    };
    auto fst4w = [&](int nzhsym, int period) {
        p.nzhsym = nzhsym; p.nmode = 241; p.ntol = 100; p.ntrperiod = period; p.nfqso = 1500; p.nexp_decode = 256 * 3;
    };
    if (m == "FT8") { p.lft8apon = true; p.nzhsym = 0; p.nmode = 8; p.napwid = 50; p.ntrperiod = 15; }
    else if (m == "FT4") { p.nmode = 5; p.ntrperiod = static_cast<int>(7.5); p.napwid = 80; p.nzhsym = 0; }
    else if (m == "Q65-30") { p.nmode = 66; p.ntxmode = 66; p.ntrperiod = 30; p.nzhsym = 196; }
    else if (m == "JT65") { p.nzhsym = 174; p.ntxmode = 65; p.nmode = 65; p.ntrperiod = 60; }
    else if (m == "FST4-60") fst4(900, 187, 60);
    else if (m == "FST4-120") fst4(900, 387, 120);
    else if (m == "FST4-300") fst4(700, 1003, 300);
    else if (m == "FST4-900") fst4(900, 3107, 900);
    else if (m == "FST4-1800") fst4(900, 6232, 1800);
    else if (m == "FST4W-120") fst4w(387, 120);
    else if (m == "FST4W-300") fst4w(1003, 300);
    else if (m == "FST4W-900") fst4w(3107, 900);
    else if (m == "FST4W-1800") fst4w(6232, 1800);
    else return false;
    dec_data->ipc[0] = p.nzhsym;
    dec_data->ipc[1] = 1;   // istart
    dec_data->ipc[2] = -1;  // idone
    std::size_t nel = item.audio.size();
    if (nel > static_cast<std::size_t>(NTMAX) * RX_SAMPLE_RATE) nel = static_cast<std::size_t>(NTMAX) * RX_SAMPLE_RATE;
    std::memcpy(&dec_data->d2[0], item.audio.data(), nel * sizeof(std::int16_t));
    return true;
}

// POSIX stand-in for QSharedMemory as the reference uses it (create(size) / attach by key / data() / detach).
class Jt9ShmSegment {
public:
    Jt9ShmSegment() = default;
    Jt9ShmSegment(const Jt9ShmSegment&) = delete;
    Jt9ShmSegment& operator=(const Jt9ShmSegment&) = delete;
    ~Jt9ShmSegment() { detach(); }
    bool create(const std::string& key) { return open(key, true); }   // mem_jt9.create(sizeof(dec_data_t))
    bool attach(const std::string& key) { return open(key, false); }  // what jt9 does with "-s <key>"
    dec_data_t* data() const { return base; }
    void detach() {
        if (base) munmap(base, sizeof(dec_data_t));
        if (fd >= 0) close(fd);
        if (owner) shm_unlink(name.c_str());
        base = nullptr;
        fd = -1;
        owner = false;
    }

private:
    bool open(const std::string& key, bool creat) {
        detach();
        name = "/" + key;
        fd = shm_open(name.c_str(), creat ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
        if (fd < 0) return false;
        owner = creat;
        if (creat && ftruncate(fd, sizeof(dec_data_t)) != 0) {
            detach();
            return false;
        }
        void* p = mmap(nullptr, sizeof(dec_data_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (p == MAP_FAILED) {
            detach();
            return false;
        }
        base = static_cast<dec_data_t*>(p);
        return true;
    }
    std::string name;
    int fd = -1;
    dec_data_t* base = nullptr;
    bool owner = false;
};
