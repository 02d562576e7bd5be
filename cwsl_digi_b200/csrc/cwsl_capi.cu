// C ABI of the B200 front-end (include/cwsl_b200.h): receiver handles, slot groups, IQ ring,
// phase-table cache and launch orchestration. No compute happens on the host: every sample is
// produced by the kernels in cwsl_kernels.cu; if CUDA is unavailable the entry points fail.
#include "../../include/cwsl_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "cwsl_kernels.hpp"
#include "cwsl_tables.hpp"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(e__ == cudaErrorMemoryAllocation ? CWSL_ERR_NOMEM : CWSL_ERR_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e__));                                                       \
    } while (0)

// ---- per-device shared state: constant taps + phase-table cache ---------------------------------
struct PhaseKey {
    int device;
    uint32_t fs;
    int32_t freq;
    int usb;
    uint32_t length;
    bool operator<(const PhaseKey& o) const {
        return std::tie(device, fs, freq, usb, length) < std::tie(o.device, o.fs, o.freq, o.usb, o.length);
    }
};
struct PhaseEntry {
    float2* table = nullptr;
    int refs = 0;
};
std::mutex g_mu;  // phase-table / anchor / per-device constant caches
std::map<PhaseKey, PhaseEntry> g_phase_cache;
std::set<std::pair<int, uint32_t>> g_taps_uploaded;
struct ChanDeviceTables {  // STFT channelizer: deconvolved window and FFT twiddles, one copy per device
    float* window = nullptr;
    float2* twiddle = nullptr;
};
std::map<std::pair<int, uint32_t>, ChanDeviceTables> g_chan_tables;  // per (device, block size)
// STFT anchor tables P_c[128 a], [anchor][channel]: shared by every slot group with the same channel list
// (64 receivers of the stress sweep share one 15 MB table instead of holding 1 GB of copies)
struct AnchorEntry {
    float2* table = nullptr;
    uint32_t n_anchor = 0;
    int refs = 0;
};
std::map<std::vector<PhaseKey>, AnchorEntry> g_anchor_cache;
// Scratch of the STFT guard's indirect FAST launches (phases replayed from the anchors, cwsl_kernels.hpp FastIndirect):
// one per (device, stream) -- launches on one stream are serialised, so every receiver queued on it shares the buffer.
struct ScratchEntry {
    float2* p = nullptr;
    int refs = 0;
};
std::map<std::pair<int, cudaStream_t>, ScratchEntry> g_fast_scratch;

// Managed hand-off buffers (cwsl_host_alloc): pinned, zero-initialised, and only ever written by
// cwsl_rx_end_slot, so the library knows which columns of a destination can hold non-zero data and
// copies only those (the zero tail of a slot buffer does not cross PCIe again).
struct OutState {
    size_t af_size = 0, rows = 0;
    size_t dirty_cols = 0;  // columns [0, dirty_cols) of every row may be non-zero on the host
};
struct HostRegion {
    size_t bytes = 0;
    std::map<uintptr_t, OutState> outs;
};
std::mutex g_host_mu;  // its own lock: a long table build under g_mu must not stall another receiver's end_slot
std::map<uintptr_t, HostRegion> g_host_regions;  // base address -> region

struct ChannelHost {
    int32_t demod_freq = 0;
    int usb = 1;
    float scale = 1.0f;
    cwsl::NcoTables nco;
};

struct Group {
    double period = 0;
    size_t af_size = 0;
    size_t af_stride = 0;
    std::vector<ChannelHost> ch;
    std::vector<PhaseKey> phase_keys;
    // channel-set changes requested while a slot is open: applied at the slot edge (cwsl_rx_end_slot)
    std::vector<ChannelHost> pending_add;
    std::vector<int> pending_remove;
    // device
    float2* d_tone = nullptr;
    const float2** d_phase = nullptr;   // [C] -> full phase tables; entries are filled by ensure_phase_tables() (EXACT / FAST launches)
    bool have_phase_tables = false;
    float2* d_pinc = nullptr;           // [C] the reference's float phase_inc per channel
    float* d_sign = nullptr;
    float* d_scale = nullptr;
    unsigned* d_maxbits = nullptr;
    float* d_factor = nullptr;
    float* d_maxval = nullptr;
    float* d_audio = nullptr;
    int16_t* d_out = nullptr;
    cwsl::ChanItem* d_chan = nullptr;   // STFT channelizer work items (<= 4 neighbouring channels each)
    uint32_t n_chan_items = 0;
    // STFT mode only, allocated at its first use (ensure_stft): anchors of the exact phase recurrence and the
    // dynamic-range guard's per-launch scratch
    const float2* d_anchors = nullptr;  // shared, see g_anchor_cache
    bool have_anchor_ref = false;
    float* d_seg_scale = nullptr;
    unsigned* d_seg_max = nullptr;
    unsigned* d_seg_energy = nullptr;
    uint32_t* d_sel = nullptr;
    cwsl::GuardItem* d_items = nullptr;
    unsigned* d_n_items = nullptr;
    unsigned long long* d_counters = nullptr;  // [0,1] this slot, [2,3] last finished slot
    uint32_t max_segs = 0;
    uint32_t tiles = 1;        // FAST tiles per segment: a function of the group's size only
    int mode = CWSL_MODE_FAST; // arithmetic mode of the OPEN slot (latched from the receiver at the slot's first launch)
    // slot state (units: SSBD blocks unless noted)
    uint64_t slot_start = 0;   // absolute index of the slot's block 0
    uint64_t processed = 0;    // slot-relative blocks already demodulated
    uint64_t iq_blocks = 0;    // IQ blocks pushed since the slot edge
    size_t last_write_index = 0;
    size_t out_dirty = 0;      // columns [0, out_dirty) of d_out may be non-zero; the quantise pass rewrites only those
    size_t out_pitch = 0;      // row pitch (samples) of the last finished slot in d_out: af_size, or write_index (packed)
    bool have_result = false;
    bool committed = false;
};

}  // namespace

struct cwsl_rx {
    int device = 0;
    ChanDeviceTables chan_tables;
    uint32_t fs = 0, iq_len = 0;
    cwsl::SsbdGeometry geo;
    uint32_t sub = 0;  // SSBD blocks per IQ block
    int mode = CWSL_MODE_FAST;
    double ring_seconds = 0;
    cudaStream_t stream = nullptr;       // kernels, H2D into the ring
    bool own_stream = true;
    cudaStream_t copy_stream = nullptr;  // D2H of finished slots, so they overlap the next kernels on `stream`
    cudaEvent_t ev_out_ready = nullptr;  // recorded on `stream` after quantise of the last slot
    cudaEvent_t ev_d2h_done = nullptr;   // recorded on `copy_stream` after the last slot's D2H
    bool d2h_pending = false;
    float2* d_ring = nullptr;        // owned ring (nullptr while bound to external IQ)
    const float2* ring_ptr = nullptr;  // what the kernels read
    uint32_t ring_blocks = 0;
    uint32_t own_ring_blocks = 0;
    uint64_t max_slot_blocks = 0;
    uint64_t abs_written = 0;  // SSBD blocks pushed since creation
    bool bound = false;
    bool committed = false;
    std::vector<Group> groups;
    // push fences: events recorded on `stream` behind pushes, so a caller that stages IQ in a ring of pinned
    // buffers can wait for exactly the copy that last read the buffer it wants to refill
    static constexpr uint64_t kFences = 64;
    cudaEvent_t fence_ev[kFences] = {};
    uint64_t fence_next = 1;  // token of the next fence; token t lives in fence_ev[t % kFences]
    float2* fast_scratch = nullptr;          // see g_fast_scratch; referenced for `fast_scratch_stream`
    cudaStream_t fast_scratch_stream = nullptr;
    double guard_db = 0;  // STFT dynamic-range guard threshold (dB below the band's mean power); <= 0: off
    // timing
    bool timing = false;
    struct DemodEvents {
        cudaEvent_t e0 = nullptr, e_pre = nullptr, e_main = nullptr, e1 = nullptr;  // e_pre/e_main: STFT launches only
    };
    std::vector<DemodEvents> ev_demod;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_quant;
    std::vector<cudaEvent_t> ev_pool;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

cudaEvent_t get_event(cwsl_rx* rx) {
    if (!rx->ev_pool.empty()) {
        cudaEvent_t e = rx->ev_pool.back();
        rx->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void release_anchor_ref(Group& g) {  // g_mu held
    if (!g.have_anchor_ref) return;
    auto it = g_anchor_cache.find(g.phase_keys);
    if (it != g_anchor_cache.end() && --it->second.refs <= 0) {
        cudaFree(it->second.table);
        g_anchor_cache.erase(it);
    }
    g.have_anchor_ref = false;
    g.d_anchors = nullptr;
}

void free_group_device(Group& g) {
    cudaFree(g.d_tone);
    cudaFree((void*)g.d_phase);
    cudaFree(g.d_pinc);
    cudaFree(g.d_sign);
    cudaFree(g.d_scale);
    cudaFree(g.d_maxbits);
    cudaFree(g.d_factor);
    cudaFree(g.d_maxval);
    cudaFree(g.d_audio);
    cudaFree(g.d_out);
    cudaFree(g.d_chan);
    cudaFree(g.d_seg_scale);
    cudaFree(g.d_seg_max);
    cudaFree(g.d_seg_energy);
    cudaFree(g.d_sel);
    cudaFree(g.d_items);
    cudaFree(g.d_n_items);
    cudaFree(g.d_counters);
    g.d_chan = nullptr;
    g.d_tone = nullptr;
    g.d_phase = nullptr;
    g.d_pinc = nullptr;
    g.d_sign = g.d_scale = g.d_factor = g.d_maxval = g.d_audio = g.d_seg_scale = nullptr;
    g.d_maxbits = g.d_seg_max = g.d_seg_energy = g.d_n_items = nullptr;
    g.d_sel = nullptr;
    g.d_items = nullptr;
    g.d_counters = nullptr;
    g.d_out = nullptr;
    g.max_segs = 0;
    g.committed = false;
    std::lock_guard<std::mutex> lk(g_mu);
    release_anchor_ref(g);
    if (g.have_phase_tables)
        for (const PhaseKey& k : g.phase_keys) {
            auto it = g_phase_cache.find(k);
            if (it != g_phase_cache.end() && --it->second.refs <= 0) {
                cudaFree(it->second.table);
                g_phase_cache.erase(it);
            }
        }
    g.have_phase_tables = false;
    g.phase_keys.clear();
}

bool commit_nosync() {
    static const bool nosync = [] {
        const char* e = std::getenv("CWSL_COMMIT_NOSYNC");
        return e && e[0] == '1';
    }();
    return nosync;
}

int commit_device_tables(cwsl_rx* rx);
int commit_group_impl(cwsl_rx* rx, Group& g);

// Upload one slot group's tables, build missing phase tables, allocate its audio buffers. On failure everything this
// attempt allocated is released again, so a later call starts from scratch instead of leaking.
int commit_group(cwsl_rx* rx, Group& g) {
    if (g.committed) return CWSL_OK;
    // One-time setup allocates GBs of device memory. Do it on a quiet device: with more than ~8 receivers on private
    // non-blocking streams, cudaMalloc overlapping other receivers' running kernels ended in "illegal memory access"
    // on driver 580 / CUDA 12.9 (any kernel, TMA or not; compute-sanitizer clean).
    // CWSL_COMMIT_NOSYNC=1 disables the synchronisation (diagnostics only).
    if (!commit_nosync()) CK(cudaDeviceSynchronize());
    const int rc = commit_group_impl(rx, g);
    if (rc != CWSL_OK) {
        const std::string msg = g_last_error;  // free_group_device must not clobber the reason
        free_group_device(g);
        g_last_error = msg;
    }
    return rc;
}

int commit(cwsl_rx* rx) {
    if (rx->committed) return CWSL_OK;
    if (rx->groups.empty()) return fail(CWSL_ERR_STATE, "no slot group defined");
    int rc = commit_device_tables(rx);
    if (rc != CWSL_OK) return rc;
    uint64_t max_blocks = 0;
    for (Group& g : rx->groups) {
        if (g.ch.empty()) return fail(CWSL_ERR_STATE, "slot group without channels");
        if ((rc = commit_group(rx, g)) != CWSL_OK) {
            const std::string msg = g_last_error;
            for (Group& h : rx->groups) free_group_device(h);
            g_last_error = msg;
            return rc;
        }
        max_blocks = std::max<uint64_t>(max_blocks, g.af_size);
    }
    rx->max_slot_blocks = max_blocks;
    rx->committed = true;
    return CWSL_OK;
}

// constant-bank taps (once per device and block size), cross-checked against the baked copy; STFT window/twiddles
int commit_device_tables(cwsl_rx* rx) {
    std::lock_guard<std::mutex> lk(g_mu);
    const auto key = std::make_pair(rx->device, rx->geo.block_size);
    if (!g_taps_uploaded.count(key)) {
        const std::vector<float> h = cwsl::lowpass_taps(rx->geo);
        const float* baked = cwsl::baked_taps_transposed(rx->geo.block_size);
        if (!baked) return fail(CWSL_ERR_INVALID, "unsupported block size %u", rx->geo.block_size);
        for (uint32_t m = 0; m < rx->geo.block_size; ++m)
            for (uint32_t n = 0; n < 32; ++n)
                if (std::memcmp(&baked[m * 32 + n], &h[rx->geo.block_size * n + m], sizeof(float)) != 0)
                    return fail(CWSL_ERR_STATE,
                                "baked low-pass taps differ from this host's libm result at tap %u "
                                "(rebuild: make -C cwsl_digi_b200/csrc taps all)",
                                rx->geo.block_size * n + m);
        CK(cwsl::upload_taps(rx->geo.block_size, h.data()));
        g_taps_uploaded.insert(key);
    }
    if (!g_chan_tables.count(key)) {  // STFT window + FFT twiddles, once per device and rate
        const std::vector<float> w = cwsl::chan_window(rx->geo, cwsl::kChanKernelWidth);
        const std::vector<std::complex<float>> tw = cwsl::chan_twiddles(rx->geo);
        ChanDeviceTables t;
        cudaError_t err = cudaMalloc(&t.window, w.size() * sizeof(float));
        if (err == cudaSuccess) err = cudaMalloc(&t.twiddle, tw.size() * sizeof(float2));
        if (err == cudaSuccess) err = cudaMemcpy(t.window, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (err == cudaSuccess) err = cudaMemcpy(t.twiddle, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice);
        if (err != cudaSuccess) {  // published only when complete
            cudaFree(t.window), cudaFree(t.twiddle);
            return fail(err == cudaErrorMemoryAllocation ? CWSL_ERR_NOMEM : CWSL_ERR_CUDA, "STFT tables: %s", cudaGetErrorString(err));
        }
        g_chan_tables[key] = t;
    }
    rx->chan_tables = g_chan_tables[key];
    return CWSL_OK;
}

int commit_group_impl(cwsl_rx* rx, Group& g) {
    const uint32_t C = (uint32_t)g.ch.size();
    if (C == 0) return fail(CWSL_ERR_STATE, "slot group without channels");
    const uint32_t BS = rx->geo.block_size;
    g.af_stride = (g.af_size + 7) / 8 * 8;
    g.tiles = cwsl::fast_tiles_per_seg(C);
    std::vector<float2> tone((size_t)C * BS), pinc(C);
    std::vector<float> sign(C), scale(C);
    for (uint32_t c = 0; c < C; ++c) {
        pinc[c] = make_float2(g.ch[c].nco.phase_inc.real(), g.ch[c].nco.phase_inc.imag());
        for (uint32_t m = 0; m < BS; ++m)
            tone[(size_t)c * BS + m] = make_float2(g.ch[c].nco.tone[m].real(), g.ch[c].nco.tone[m].imag());
        sign[c] = g.ch[c].nco.sign;
        scale[c] = g.ch[c].scale;
    }
    CK(cudaMalloc(&g.d_tone, tone.size() * sizeof(float2)));
    CK(cudaMalloc((void**)&g.d_phase, C * sizeof(float2*)));
    CK(cudaMemsetAsync((void*)g.d_phase, 0, C * sizeof(float2*), rx->stream));
    CK(cudaMalloc(&g.d_pinc, C * sizeof(float2)));
    CK(cudaMalloc(&g.d_sign, C * sizeof(float)));
    CK(cudaMalloc(&g.d_scale, C * sizeof(float)));
    CK(cudaMalloc(&g.d_maxbits, C * sizeof(unsigned)));
    CK(cudaMalloc(&g.d_factor, C * sizeof(float)));
    CK(cudaMalloc(&g.d_maxval, C * sizeof(float)));
    CK(cudaMalloc(&g.d_audio, (size_t)C * g.af_stride * sizeof(float)));
    CK(cudaMalloc(&g.d_out, (size_t)C * g.af_size * sizeof(int16_t)));
    g.out_dirty = g.af_size;  // nothing is known about the fresh buffer: the first slot writes every column
    CK(cudaMalloc(&g.d_counters, 4 * sizeof(unsigned long long)));
    CK(cudaMemcpyAsync(g.d_tone, tone.data(), tone.size() * sizeof(float2), cudaMemcpyHostToDevice, rx->stream));
    CK(cudaMemcpyAsync(g.d_pinc, pinc.data(), C * sizeof(float2), cudaMemcpyHostToDevice, rx->stream));
    CK(cudaMemcpyAsync(g.d_sign, sign.data(), C * sizeof(float), cudaMemcpyHostToDevice, rx->stream));
    CK(cudaMemcpyAsync(g.d_scale, scale.data(), C * sizeof(float), cudaMemcpyHostToDevice, rx->stream));
    CK(cudaMemsetAsync(g.d_maxbits, 0, C * sizeof(unsigned), rx->stream));
    CK(cudaMemsetAsync(g.d_counters, 0, 4 * sizeof(unsigned long long), rx->stream));
    // STFT channelizer work items (CWSL_MODE_STFT): channels in ascending grid position, <= 4 neighbours per item
    std::vector<cwsl::NcoTables> ncos(C);
    for (uint32_t c = 0; c < C; ++c) ncos[c] = g.ch[c].nco;
    const std::vector<cwsl::ChanItemHost> hitems =
        cwsl::chan_item_order(cwsl::chan_items(rx->geo, ncos, cwsl::kChanKernelWidth), (int)cwsl::chan_grid(rx->geo), cwsl::kChanMaxItems);
    std::vector<cwsl::ChanItem> items(hitems.size());
    const int grid_bins = (int)cwsl::chan_grid(rx->geo);
    for (size_t i = 0; i < hitems.size(); ++i) {
        const cwsl::ChanItemHost& h = hitems[i];
        cwsl::ChanItem& it = items[i];
        std::memset(&it, 0, sizeof(it));
        it.bin_off = (uint32_t)(((h.e % grid_bins) + grid_bins) % grid_bins) * 8u;
        it.n_members = (uint32_t)h.n;
        for (int j = 0; j < cwsl::kChanItemMembers; ++j) {
            it.ch[j] = h.ch[j] < 0 ? 0xffffffffu : (uint32_t)h.ch[j];
            if (h.ch[j] < 0) continue;
            const ChannelHost& ch = g.ch[h.ch[j]];
            it.sign[j] = ch.nco.sign;
            it.pinc[j][0] = ch.nco.phase_inc.real();
            it.pinc[j][1] = ch.nco.phase_inc.imag();
            it.rot[j][0] = h.rot[j].real();
            it.rot[j][1] = h.rot[j].imag();
            for (int q = 0; q < cwsl::kChanItemTaps; ++q) it.w[j][q] = h.w[j][q];
        }
    }
    g.n_chan_items = (uint32_t)items.size();
    CK(cudaMalloc(&g.d_chan, items.size() * sizeof(cwsl::ChanItem)));
    CK(cudaMemcpyAsync(g.d_chan, items.data(), items.size() * sizeof(cwsl::ChanItem), cudaMemcpyHostToDevice, rx->stream));
    CK(cudaStreamSynchronize(rx->stream));  // host vectors go out of scope

    // keys of the channels' phase recurrences (tables and anchors are built when a launch first needs them)
    const uint32_t length = (uint32_t)((g.af_size + 3) / 4 * 4 + 4);
    g.phase_keys.clear();
    for (uint32_t c = 0; c < C; ++c) g.phase_keys.push_back(PhaseKey{rx->device, rx->fs, g.ch[c].demod_freq, g.ch[c].usb, length});
    g.have_phase_tables = false;
    g.committed = true;
    return CWSL_OK;
}

// Full phase tables P_c[k] = phase_inc_c^k (8 bytes per audio sample and channel), needed by the EXACT kernel and by
// direct FAST launches; the STFT mode gets by with the anchors (ensure_stft) and never calls this. One table per
// distinct (Fs, demodFreq, sideband, length), shared process-wide. New tables are built into local pointers and
// PUBLISHED to the cache only when they are complete: a failed build leaves nothing half-made behind, and nobody can
// ever pick up a table that is allocated but not yet filled.
int ensure_phase_tables(cwsl_rx* rx, Group& g) {
    if (g.have_phase_tables) return CWSL_OK;
    const uint32_t C = (uint32_t)g.ch.size();
    if (!commit_nosync()) CK(cudaDeviceSynchronize());  // one-time allocations on a quiet device, see commit_group()
    std::vector<const float2*> ptrs(C);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        std::map<PhaseKey, float2*> fresh;
        std::vector<float2> new_inc;
        std::vector<float2*> new_tab;
        auto drop_fresh = [&] {
            for (auto& kv : fresh) cudaFree(kv.second);
        };
        const uint32_t length = C ? g.phase_keys[0].length : 0;
        for (uint32_t c = 0; c < C; ++c) {
            const PhaseKey& k = g.phase_keys[c];
            if (g_phase_cache.count(k) || fresh.count(k)) continue;
            float2* tab = nullptr;
            const cudaError_t err = cudaMalloc(&tab, (size_t)length * sizeof(float2));
            if (err != cudaSuccess) {
                drop_fresh();
                return fail(CWSL_ERR_NOMEM, "phase table alloc: %s", cudaGetErrorString(err));
            }
            fresh[k] = tab;
            new_inc.push_back(make_float2(g.ch[c].nco.phase_inc.real(), g.ch[c].nco.phase_inc.imag()));
            new_tab.push_back(tab);
        }
        if (!new_tab.empty()) {
            float2* d_inc = nullptr;
            float2** d_tab = nullptr;
            cudaError_t err = cudaMalloc(&d_inc, new_inc.size() * sizeof(float2));
            if (err == cudaSuccess) err = cudaMalloc((void**)&d_tab, new_tab.size() * sizeof(float2*));
            if (err == cudaSuccess) err = cudaMemcpy(d_inc, new_inc.data(), new_inc.size() * sizeof(float2), cudaMemcpyHostToDevice);
            if (err == cudaSuccess) err = cudaMemcpy((void*)d_tab, new_tab.data(), new_tab.size() * sizeof(float2*), cudaMemcpyHostToDevice);
            if (err == cudaSuccess) err = cwsl::launch_phase_tables(d_inc, d_tab, (uint32_t)new_tab.size(), length, rx->stream);
            if (err == cudaSuccess) err = cudaStreamSynchronize(rx->stream);
            cudaFree(d_inc);
            cudaFree((void*)d_tab);
            if (err != cudaSuccess) {
                drop_fresh();
                return fail(err == cudaErrorMemoryAllocation ? CWSL_ERR_NOMEM : CWSL_ERR_CUDA, "phase table build: %s",
                            cudaGetErrorString(err));
            }
        }
        for (auto& kv : fresh) g_phase_cache[kv.first].table = kv.second;  // publish
        for (uint32_t c = 0; c < C; ++c) {
            PhaseEntry& e = g_phase_cache[g.phase_keys[c]];
            ++e.refs;
            ptrs[c] = e.table;
        }
        g.have_phase_tables = true;
    }
    CK(cudaMemcpyAsync((void*)g.d_phase, ptrs.data(), C * sizeof(float2*), cudaMemcpyHostToDevice, rx->stream));
    CK(cudaStreamSynchronize(rx->stream));  // (ptrs goes out of scope)
    return CWSL_OK;
}

void release_fast_scratch(cwsl_rx* rx) {
    if (!rx->fast_scratch) return;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_fast_scratch.find({rx->device, rx->fast_scratch_stream});
    if (it != g_fast_scratch.end() && --it->second.refs <= 0) {
        cudaFree(it->second.p);
        g_fast_scratch.erase(it);
    }
    rx->fast_scratch = nullptr;
    rx->fast_scratch_stream = nullptr;
}

int ensure_fast_scratch(cwsl_rx* rx) {
    if (rx->fast_scratch && rx->fast_scratch_stream == rx->stream) return CWSL_OK;
    release_fast_scratch(rx);
    std::lock_guard<std::mutex> lk(g_mu);
    ScratchEntry& e = g_fast_scratch[{rx->device, rx->stream}];
    if (!e.p) {
        cudaError_t err = commit_nosync() ? cudaSuccess : cudaDeviceSynchronize();
        if (err == cudaSuccess) err = cudaMalloc(&e.p, cwsl::fast_scratch_bytes(rx->device));
        if (err != cudaSuccess) {
            g_fast_scratch.erase({rx->device, rx->stream});
            return fail(err == cudaErrorMemoryAllocation ? CWSL_ERR_NOMEM : CWSL_ERR_CUDA, "guard scratch: %s", cudaGetErrorString(err));
        }
    }
    ++e.refs;
    rx->fast_scratch = e.p;
    rx->fast_scratch_stream = rx->stream;
    return CWSL_OK;
}

// STFT-mode extras of a slot group, set up at its first STFT launch: the anchor table of the exact phase recurrence
// (shared between groups with the same channel list) and the guard's scratch.
int ensure_stft(cwsl_rx* rx, Group& g) {
    const uint32_t C = (uint32_t)g.ch.size();
    if (!g.d_anchors) {
        std::lock_guard<std::mutex> lk(g_mu);
        AnchorEntry& e = g_anchor_cache[g.phase_keys];
        if (!e.table) {
            const uint32_t n_anchor = (uint32_t)(g.af_size / cwsl::kChanAnchorHops + 1);  // 128 a <= af_size < table length
            float2* tab = nullptr;
            cudaError_t err = commit_nosync() ? cudaSuccess : cudaDeviceSynchronize();
            if (err == cudaSuccess) err = cudaMalloc(&tab, (size_t)n_anchor * C * sizeof(float2));
            if (err == cudaSuccess) err = cwsl::launch_phase_anchors(g.d_pinc, tab, C, n_anchor, cwsl::kChanAnchorHops, rx->stream);
            if (err == cudaSuccess) err = cudaStreamSynchronize(rx->stream);
            if (err != cudaSuccess) {
                cudaFree(tab);
                g_anchor_cache.erase(g.phase_keys);
                return fail(err == cudaErrorMemoryAllocation ? CWSL_ERR_NOMEM : CWSL_ERR_CUDA, "anchor table: %s",
                            cudaGetErrorString(err));
            }
            e.table = tab;
            e.n_anchor = n_anchor;
        }
        ++e.refs;
        g.have_anchor_ref = true;
        g.d_anchors = e.table;
    }
    if (g.max_segs == 0) {  // (set last: a failed attempt is rolled back and repeated by the next launch)
        const uint32_t W = cwsl::fast_seg_blocks(g.tiles);
        const uint32_t segs = (uint32_t)(g.af_size / W + 2);
        if (!commit_nosync()) CK(cudaDeviceSynchronize());
        cudaError_t err = cudaMalloc(&g.d_seg_scale, segs * sizeof(float));
        if (err == cudaSuccess) err = cudaMalloc(&g.d_seg_max, (size_t)segs * C * sizeof(unsigned));
        if (err == cudaSuccess) err = cudaMalloc(&g.d_seg_energy, (size_t)segs * C * sizeof(unsigned));
        if (err == cudaSuccess) err = cudaMalloc(&g.d_sel, (size_t)segs * C * sizeof(uint32_t));
        if (err == cudaSuccess) err = cudaMalloc(&g.d_items, (size_t)segs * ((C + 31) / 32) * sizeof(cwsl::GuardItem));
        if (err == cudaSuccess) err = cudaMalloc(&g.d_n_items, sizeof(unsigned));
        if (err == cudaSuccess) err = cudaMemsetAsync(g.d_seg_max, 0, (size_t)segs * C * sizeof(unsigned), rx->stream);
        if (err == cudaSuccess) err = cudaMemsetAsync(g.d_seg_energy, 0, (size_t)segs * C * sizeof(unsigned), rx->stream);
        if (err == cudaSuccess) err = cudaMemsetAsync(g.d_n_items, 0, sizeof(unsigned), rx->stream);
        if (err != cudaSuccess) {
            cudaFree(g.d_seg_scale), cudaFree(g.d_seg_max), cudaFree(g.d_seg_energy);
            cudaFree(g.d_sel), cudaFree(g.d_items), cudaFree(g.d_n_items);
            g.d_seg_scale = nullptr, g.d_seg_max = g.d_seg_energy = g.d_n_items = nullptr, g.d_sel = nullptr, g.d_items = nullptr;
            cudaGetLastError();
            return fail(err == cudaErrorMemoryAllocation ? CWSL_ERR_NOMEM : CWSL_ERR_CUDA, "STFT guard scratch: %s",
                        cudaGetErrorString(err));
        }
        g.max_segs = segs;
    }
    return CWSL_OK;
}

// The receiver's own device ring, allocated on the first host/device push (a receiver that is only
// ever bound to caller-resident IQ never needs one). Leaves "bound" mode: the next slot starts empty.
int ensure_ring(cwsl_rx* rx) {
    if (!rx->d_ring) {
        const double secs = rx->ring_seconds;
        uint64_t blocks;
        if (secs > 0)
            blocks = (uint64_t)(secs * rx->fs / rx->geo.block_size);
        else
            blocks = rx->max_slot_blocks + 64;  // longest slot (+5 s) fits without intermediate demodulation
        const uint64_t quantum = std::max<uint64_t>(rx->sub, 4);
        // The FAST / STFT modes demodulate whole segments only (fixed slot-relative positions, so that the result
        // does not depend on the chunking); up to one segment (+31 blocks of history) of every open slot stays in
        // the ring while the next push, at most half a ring long, arrives. ring_seconds is therefore a minimum.
        uint64_t seg = 0;
        for (const Group& g : rx->groups) seg = std::max<uint64_t>(seg, cwsl::fast_seg_blocks(g.tiles));
        blocks = std::max<uint64_t>(blocks, 2 * seg + 4 * (uint64_t)rx->sub + 64);
        blocks = (blocks + quantum - 1) / quantum * quantum;
        rx->own_ring_blocks = (uint32_t)blocks;
        CK(cudaDeviceSynchronize());  // allocate on a quiet device, see commit()
        CK(cudaMalloc(&rx->d_ring, (size_t)blocks * rx->geo.block_size * sizeof(float2)));
    }
    if (rx->bound || rx->ring_ptr != rx->d_ring) {
        rx->bound = false;
        rx->ring_ptr = rx->d_ring;
        rx->ring_blocks = rx->own_ring_blocks;
        rx->abs_written = 0;
        for (Group& g : rx->groups) {
            g.slot_start = 0;
            g.processed = 0;
            g.iq_blocks = 0;
        }
    }
    return CWSL_OK;
}

// SSBD blocks of the current slot that are accepted by the reference's af-buffer guard.
uint64_t slot_target_blocks(const cwsl_rx* rx, const Group& g) {
    const size_t acc = cwsl::accepted_blocks((size_t)g.iq_blocks, rx->iq_len, rx->geo.dec_ratio, g.af_size);
    return (uint64_t)acc * rx->sub;
}

// CWSL_MODE_STFT uses the channelizer from this many channels per slot group on. Measured break-even on B200 at
// 192 kHz: 64 channels x FT8 slot take 0.39 ms in the direct FAST kernel and 0.34 ms in the channelizer (whose FFT
// side alone is 0.33 ms); 48 channels 0.29 vs 0.33 ms; 256 channels 1.27 vs 0.38 ms. CWSL_STFT_MIN_CHANNELS overrides.
uint32_t stft_min_channels() {
    static const uint32_t v = [] {
        const char* e = std::getenv("CWSL_STFT_MIN_CHANNELS");
        return e ? (uint32_t)std::max(1, std::atoi(e)) : 64u;
    }();
    return v;
}

// Default threshold of the STFT dynamic-range guard, dB below the band's mean power (cwsl_guard.cu). Measured float32
// FFT floor: at worst 5.8e-7 of the band's rms (-124.7 dB; single dominant carrier), typically 1.9e-7; 90 dB above
// the worst case is -34.7 dB, the default leaves 2.7 dB of margin. (The stress sweep's channels sit at -18 ... -24 dB.)
// CWSL_STFT_GUARD_DB overrides (0 = guard off).
double default_guard_db() {
    static const double v = [] {
        const char* e = std::getenv("CWSL_STFT_GUARD_DB");
        return e ? std::atof(e) : 32.0;
    }();
    return v;
}

// Demodulate what has been pushed for the group's open slot. EXACT: everything. FAST / STFT: whole segments only
// unless `final` (the slot edge), so every launch starts on a segment boundary.
int process_group(cwsl_rx* rx, Group& g, bool final) {
    uint64_t target = slot_target_blocks(rx, g);
    if (g.processed == 0) g.mode = rx->mode;  // the arithmetic mode is latched per slot
    const uint32_t seg = g.mode == CWSL_MODE_EXACT ? 0u : cwsl::fast_seg_blocks(g.tiles);
    if (seg && !final) target = target / seg * seg;
    if (target <= g.processed) return CWSL_OK;
    if (seg && g.processed % seg != 0) return fail(CWSL_ERR_STATE, "internal: launch not on a segment boundary");
    // history still in the ring?
    const uint64_t first_needed = g.slot_start + (g.processed >= 31 ? g.processed - 31 : 0);
    if (rx->abs_written - first_needed > rx->ring_blocks)
        return fail(CWSL_ERR_OVERRUN, "IQ ring overrun: slot needs block %llu, ring holds the last %u",
                    (unsigned long long)first_needed, rx->ring_blocks);
    cwsl::DemodLaunch p;
    p.iq_ring = rx->ring_ptr;
    p.ring_blocks = rx->ring_blocks;
    p.ring_off = (uint32_t)(g.slot_start % rx->ring_blocks);
    p.block_size = rx->geo.block_size;
    p.b0 = (uint32_t)g.processed;
    p.b1 = (uint32_t)target;
    p.n_channels = (uint32_t)g.ch.size();
    p.tone = g.d_tone;
    p.phase = g.d_phase;
    p.sign = g.d_sign;
    p.audio = g.d_audio;
    p.af_stride = g.af_stride;
    p.maxbits = g.d_maxbits;
    const bool stft = g.mode == CWSL_MODE_STFT && p.n_channels >= stft_min_channels() && p.ring_blocks >= 64;
    // (IQ buffers shorter than two windows stay with the direct kernel)
    if (stft) {
        int rc = ensure_stft(rx, g);
        if (rc == CWSL_OK && rx->guard_db > 0) rc = ensure_fast_scratch(rx);
        if (rc != CWSL_OK) return rc;
    } else {  // EXACT and direct FAST launches read the full phase tables
        const int rc = ensure_phase_tables(rx, g);
        if (rc != CWSL_OK) return rc;
    }
    // the post stream may still be normalising / copying the previous slot out of the buffers this launch rewrites
    if (rx->d2h_pending) CK(cudaStreamWaitEvent(rx->stream, rx->ev_d2h_done, 0));
    cwsl_rx::DemodEvents ev;
    if (rx->timing) {
        ev.e0 = get_event(rx);
        ev.e1 = get_event(rx);
        CK(cudaEventRecord(ev.e0, rx->stream));
    }
    if (g.mode == CWSL_MODE_EXACT) {
        // CWSL_EXACT_KERNEL=gather selects the independent one-thread-per-output implementation (cross-check)
        static const bool gather = [] {
            const char* e = std::getenv("CWSL_EXACT_KERNEL");
            return e && std::string(e) == "gather";
        }();
        if (gather)
            CK(cwsl::launch_demod_exact_gather(p, rx->stream));
        else
            CK(cwsl::launch_demod_exact(p, rx->stream));
    } else if (stft) {
        // big channel groups: one FFT per hop shared by all channels, <= kChanMaxItems work items (of up to 4
        // neighbouring channels) per launch;
        // channel segments too close to the FFT's noise floor are redone by the direct-form kernel (cwsl_guard.cu)
        const bool guard = rx->guard_db > 0;
        cwsl::GuardLaunch gl;
        gl.seg_blocks = seg;
        gl.t2 = (float)std::pow(10.0, -rx->guard_db / 10.0);
        gl.seg_scale = g.d_seg_scale;
        gl.seg_max = g.d_seg_max;
        gl.seg_energy = g.d_seg_energy;
        gl.stat_stride = p.n_channels;
        gl.sel = g.d_sel;
        gl.sel_stride = p.n_channels;
        gl.items = g.d_items;
        gl.n_items = g.d_n_items;
        gl.counters = g.d_counters;
        if ((p.b1 - p.b0 + seg - 1) / seg > g.max_segs) return fail(CWSL_ERR_STATE, "internal: guard scratch too small");
        if (guard) CK(cwsl::launch_guard_band_power(p, gl, rx->stream));
        if (rx->timing) {
            ev.e_pre = get_event(rx);
            ev.e_main = get_event(rx);
            CK(cudaEventRecord(ev.e_pre, rx->stream));
        }
        for (uint32_t i0 = 0; i0 < g.n_chan_items; i0 += cwsl::kChanMaxItems) {
            cwsl::ChanLaunch c;
            c.window = rx->chan_tables.window;
            c.twiddle = rx->chan_tables.twiddle;
            c.items = g.d_chan + i0;
            c.n_items = std::min<uint32_t>(cwsl::kChanMaxItems, g.n_chan_items - i0);
            c.anchors = g.d_anchors;
            c.anchor_stride = p.n_channels;
            if (guard) {
                c.seg_blocks = seg;
                c.seg_max = g.d_seg_max;
                c.seg_energy = g.d_seg_energy;
                c.seg_scale = g.d_seg_scale;
                c.stat_stride = p.n_channels;
            }
            CK(cwsl::launch_demod_chan(p, c, rx->stream));
        }
        if (rx->timing) CK(cudaEventRecord(ev.e_main, rx->stream));
        if (guard) {
            CK(cwsl::launch_guard_select(p, gl, rx->stream));
            cwsl::FastIndirect ind;
            ind.items = g.d_items;
            ind.n_items = g.d_n_items;
            ind.sel = g.d_sel;
            ind.sel_stride = p.n_channels;
            ind.anchors = g.d_anchors;  // phases replayed from the exact checkpoints: no full tables in STFT mode
            ind.anchor_stride = p.n_channels;
            ind.pinc = g.d_pinc;
            ind.scratch = rx->fast_scratch;
            CK(cwsl::launch_demod_fast(p, g.tiles, ind, rx->stream));
        }
    } else {
        CK(cwsl::launch_demod_fast(p, g.tiles, cwsl::FastIndirect{}, rx->stream));
    }
    if (rx->timing) {
        CK(cudaEventRecord(ev.e1, rx->stream));
        rx->ev_demod.push_back(ev);
    }
    g.processed = target;
    return CWSL_OK;
}

int push_common(cwsl_rx* rx, const float* iq, size_t n_blocks, cudaMemcpyKind kind) {
    if (!rx) return fail(CWSL_ERR_INVALID, "null receiver");
    if (n_blocks == 0) return CWSL_OK;
    if (!iq) return fail(CWSL_ERR_INVALID, "null IQ pointer");
    DeviceGuard dg(rx->device);
    if (!dg.ok) return fail(CWSL_ERR_CUDA, "cudaSetDevice(%d) failed", rx->device);
    int rc = commit(rx);
    if (rc != CWSL_OK) return rc;
    if ((rc = ensure_ring(rx)) != CWSL_OK) return rc;
    const uint32_t BS = rx->geo.block_size;
    // never let a single copy cover more than half the ring, so open slots can be drained first
    const size_t max_chunk = std::max<size_t>(1, (rx->ring_blocks / 2) / rx->sub);
    size_t done = 0;
    while (done < n_blocks) {
        const size_t nb = std::min(max_chunk, n_blocks - done);
        const uint64_t add = (uint64_t)nb * rx->sub;
        // demodulate any slot whose un-demodulated samples (plus 31 blocks of history) would be overwritten
        for (Group& g : rx->groups) {
            const uint64_t first_needed = g.slot_start + (g.processed >= 31 ? g.processed - 31 : 0);
            if (rx->abs_written + add - first_needed > rx->ring_blocks) {
                rc = process_group(rx, g, false);
                if (rc != CWSL_OK) return rc;
            }
        }
        const uint64_t row = rx->abs_written % rx->ring_blocks;
        const uint64_t first = std::min<uint64_t>(add, rx->ring_blocks - row);
        const float* src = iq + done * (size_t)rx->iq_len * 2;
        CK(cudaMemcpyAsync(rx->d_ring + row * BS, src, first * BS * sizeof(float2), kind, rx->stream));
        if (first < add)
            CK(cudaMemcpyAsync(rx->d_ring, src + first * BS * 2, (add - first) * BS * sizeof(float2), kind, rx->stream));
        rx->abs_written += add;
        for (Group& g : rx->groups) g.iq_blocks += nb;
        done += nb;
    }
    return CWSL_OK;
}

Group* get_group(cwsl_rx* rx, int group) {
    if (!rx || group < 0 || group >= (int)rx->groups.size()) {
        fail(CWSL_ERR_INVALID, "bad receiver/group %d", group);
        return nullptr;
    }
    return &rx->groups[group];
}

}  // namespace

extern "C" {

int cwsl_abi_version(void) { return CWSL_B200_ABI_VERSION; }

const char* cwsl_last_error(void) { return g_last_error.c_str(); }

int cwsl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int cwsl_ssbd_params(uint32_t sample_rate, uint32_t out[9]) {
    cwsl::SsbdGeometry g;
    if (!cwsl::ssbd_geometry(sample_rate, &g)) return fail(CWSL_ERR_INVALID, "Fs/B must be an even integer >= 4");
    out[0] = g.fs;
    out[1] = 2 * cwsl::kSSBBW;
    out[2] = g.in_size;
    out[3] = 4;
    out[4] = cwsl::kSSBBW;
    out[5] = 1u << cwsl::kLatencyLog2;
    out[6] = g.filt_order;
    out[7] = g.block_size;
    out[8] = g.num_ws;
    return CWSL_OK;
}

int cwsl_build_tables(uint32_t sample_rate, int32_t demod_freq_hz, int is_usb, float* filter, float* tone,
                      float* phase_inc) {
    cwsl::SsbdGeometry g;
    if (!cwsl::ssbd_geometry(sample_rate, &g)) return fail(CWSL_ERR_INVALID, "Fs/B must be an even integer >= 4");
    cwsl::NcoTables t;
    if (!cwsl::nco_tables(g, demod_freq_hz, is_usb != 0, &t)) return fail(CWSL_ERR_INVALID, "Signal outside of band");
    const std::vector<float> h = cwsl::lowpass_taps(g);
    if (filter) std::memcpy(filter, h.data(), h.size() * sizeof(float));
    if (tone)
        for (uint32_t n = 0; n < g.block_size; ++n) {
            tone[2 * n] = t.tone[n].real();
            tone[2 * n + 1] = t.tone[n].imag();
        }
    if (phase_inc) {
        phase_inc[0] = t.phase_inc.real();
        phase_inc[1] = t.phase_inc.imag();
    }
    return CWSL_OK;
}

int cwsl_stft_tables(uint32_t sample_rate, float* window, float* twiddle) {
    cwsl::SsbdGeometry g;
    if (!cwsl::ssbd_geometry(sample_rate, &g)) return fail(CWSL_ERR_INVALID, "Fs/B must be an even integer >= 4");
    if (g.block_size != 16 && g.block_size != 8 && g.block_size != 4)
        return fail(CWSL_ERR_INVALID, "the STFT channelizer is built for 192, 96 and 48 kHz receivers");
    if (window) {
        const std::vector<float> w = cwsl::chan_window(g, cwsl::kChanKernelWidth);
        std::memcpy(window, w.data(), w.size() * sizeof(float));
    }
    if (twiddle) {
        const std::vector<std::complex<float>> tw = cwsl::chan_twiddles(g);
        for (size_t i = 0; i < tw.size(); ++i) {
            twiddle[2 * i] = tw[i].real();
            twiddle[2 * i + 1] = tw[i].imag();
        }
    }
    return CWSL_OK;
}

int cwsl_stft_channel(uint32_t sample_rate, int32_t demod_freq_hz, int is_usb, int32_t* q0, float* wgt, float* rot) {
    cwsl::SsbdGeometry g;
    if (!cwsl::ssbd_geometry(sample_rate, &g)) return fail(CWSL_ERR_INVALID, "Fs/B must be an even integer >= 4");
    if (g.block_size != 16 && g.block_size != 8 && g.block_size != 4)
        return fail(CWSL_ERR_INVALID, "the STFT channelizer is built for 192, 96 and 48 kHz receivers");
    cwsl::NcoTables t;
    if (!cwsl::nco_tables(g, demod_freq_hz, is_usb != 0, &t)) return fail(CWSL_ERR_INVALID, "Signal outside of band");
    const cwsl::ChanChannel c = cwsl::chan_channel(g, t, cwsl::kChanKernelWidth, cwsl::kChanTaps);
    if (q0) *q0 = c.q0;
    if (wgt) std::memcpy(wgt, c.wgt.data(), c.wgt.size() * sizeof(float));
    if (rot) {
        rot[0] = c.rot.real();
        rot[1] = c.rot.imag();
    }
    return CWSL_OK;
}

int cwsl_stft_items(uint32_t sample_rate, const int32_t* demod_freq_hz, const int* is_usb, uint32_t n, int32_t* first_bin,
                    int32_t* channels, float* weights, uint32_t* n_items) {
    cwsl::SsbdGeometry g;
    if (!cwsl::ssbd_geometry(sample_rate, &g)) return fail(CWSL_ERR_INVALID, "Fs/B must be an even integer >= 4");
    if (g.block_size != 16 && g.block_size != 8 && g.block_size != 4)
        return fail(CWSL_ERR_INVALID, "the STFT channelizer is built for 192, 96 and 48 kHz receivers");
    if (!demod_freq_hz || !n_items) return fail(CWSL_ERR_INVALID, "null argument");
    std::vector<cwsl::NcoTables> nco(n);
    for (uint32_t c = 0; c < n; ++c)
        if (!cwsl::nco_tables(g, demod_freq_hz[c], is_usb ? is_usb[c] != 0 : true, &nco[c]))
            return fail(CWSL_ERR_INVALID, "Signal outside of band");
    // in the kernel's own thread order (launches of kChanMaxItems items, one item per interpolation thread)
    const std::vector<cwsl::ChanItemHost> items =
        cwsl::chan_item_order(cwsl::chan_items(g, nco, cwsl::kChanKernelWidth), (int)cwsl::chan_grid(g), cwsl::kChanMaxItems);
    for (size_t i = 0; i < items.size(); ++i) {
        if (first_bin) first_bin[i] = items[i].e;
        for (int j = 0; j < cwsl::kItemMembers; ++j) {
            if (channels) channels[i * cwsl::kItemMembers + j] = items[i].ch[j];
            if (weights)
                for (int q = 0; q < cwsl::kItemTaps; ++q)
                    weights[(i * cwsl::kItemMembers + j) * cwsl::kItemTaps + q] = items[i].w[j][q];
        }
    }
    *n_items = (uint32_t)items.size();
    return CWSL_OK;
}

size_t cwsl_af_size(double period_s) { return cwsl::af_size(static_cast<float>(period_s)); }

size_t cwsl_accepted_blocks(size_t n_iq_blocks, uint32_t iq_len, uint32_t sample_rate, size_t af_size) {
    if (sample_rate < cwsl::kWaveSR || iq_len == 0 || af_size == 0) return 0;
    return cwsl::accepted_blocks(n_iq_blocks, iq_len, sample_rate / cwsl::kWaveSR, af_size);
}

cwsl_rx_t* cwsl_rx_create(int device, uint32_t sample_rate, uint32_t iq_len, double ring_seconds) {
    cwsl::SsbdGeometry geo;
    if (!cwsl::ssbd_geometry(sample_rate, &geo)) {
        fail(CWSL_ERR_INVALID, "Fs/B must be an even integer >= 4 (Fs=%u)", sample_rate);
        return nullptr;
    }
    if (geo.block_size != 16 && geo.block_size != 8 && geo.block_size != 4) {
        fail(CWSL_ERR_INVALID, "unsupported sample rate %u (supported: 48000, 96000, 192000)", sample_rate);
        return nullptr;
    }
    if (iq_len == 0 || iq_len % geo.in_size != 0) {
        fail(CWSL_ERR_INVALID, "iq_len %u must be a positive multiple of SSBD::GetInSize() = %u", iq_len, geo.in_size);
        return nullptr;
    }
    if (ring_seconds < 0) {
        fail(CWSL_ERR_INVALID, "ring_seconds < 0");
        return nullptr;
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        fail(CWSL_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
        return nullptr;
    }
    if (device < 0 || device >= n) {
        fail(CWSL_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
        return nullptr;
    }
    DeviceGuard dg(device);
    if (!dg.ok) {
        fail(CWSL_ERR_CUDA, "cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    std::unique_ptr<cwsl_rx> rx(new cwsl_rx);
    rx->device = device;
    rx->fs = sample_rate;
    rx->iq_len = iq_len;
    rx->geo = geo;
    rx->sub = iq_len / geo.block_size;
    rx->ring_seconds = ring_seconds;
    rx->guard_db = default_guard_db();
    // The post stream (normalise/quantise, max reset, D2H) gets the LOWEST priority: when its quantise kernel and the
    // next demodulation become runnable together, the block scheduler places the demodulator's CTAs first and the
    // HBM-bound quantise CTAs fill what is left of every SM, instead of the demodulator waiting for them to drain.
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (const char* e = std::getenv("CWSL_STREAM_PRIORITIES"); e && e[0] == '0') prio_least = prio_greatest = 0;  // diagnostics
    if (cudaStreamCreateWithPriority(&rx->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
        cudaStreamCreateWithPriority(&rx->copy_stream, cudaStreamNonBlocking, prio_least) != cudaSuccess ||
        cudaEventCreateWithFlags(&rx->ev_out_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&rx->ev_d2h_done, cudaEventDisableTiming) != cudaSuccess) {
        fail(CWSL_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (rx->stream) cudaStreamDestroy(rx->stream);  // whatever part of the set exists
        if (rx->copy_stream) cudaStreamDestroy(rx->copy_stream);
        if (rx->ev_out_ready) cudaEventDestroy(rx->ev_out_ready);
        if (rx->ev_d2h_done) cudaEventDestroy(rx->ev_d2h_done);
        return nullptr;
    }
    return rx.release();
}

void cwsl_rx_destroy(cwsl_rx_t* rx) {
    if (!rx) return;
    DeviceGuard dg(rx->device);
    if (rx->stream) cudaStreamSynchronize(rx->stream);
    if (rx->copy_stream) cudaStreamSynchronize(rx->copy_stream);
    for (Group& g : rx->groups) free_group_device(g);
    release_fast_scratch(rx);
    cudaFree(rx->d_ring);
    for (auto& ev : rx->ev_demod)
        for (cudaEvent_t e : {ev.e0, ev.e_pre, ev.e_main, ev.e1})
            if (e) cudaEventDestroy(e);
    for (auto& pr : rx->ev_quant) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    for (auto e : rx->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : rx->fence_ev)
        if (e) cudaEventDestroy(e);
    if (rx->stream && rx->own_stream) cudaStreamDestroy(rx->stream);
    if (rx->copy_stream) cudaStreamDestroy(rx->copy_stream);
    if (rx->ev_out_ready) cudaEventDestroy(rx->ev_out_ready);
    if (rx->ev_d2h_done) cudaEventDestroy(rx->ev_d2h_done);
    delete rx;
}

int cwsl_rx_set_mode(cwsl_rx_t* rx, int mode) {
    if (!rx || (mode != CWSL_MODE_EXACT && mode != CWSL_MODE_FAST && mode != CWSL_MODE_STFT)) return fail(CWSL_ERR_INVALID, "bad mode %d", mode);
    rx->mode = mode;
    return CWSL_OK;
}

int cwsl_rx_set_stream(cwsl_rx_t* rx, void* cuda_stream) {
    if (!rx) return fail(CWSL_ERR_INVALID, "null receiver");
    DeviceGuard dg(rx->device);
    CK(cudaStreamSynchronize(rx->stream));
    CK(cudaStreamSynchronize(rx->copy_stream));
    rx->d2h_pending = false;
    release_fast_scratch(rx);
    if (rx->own_stream && rx->stream) cudaStreamDestroy(rx->stream);
    rx->stream = static_cast<cudaStream_t>(cuda_stream);
    rx->own_stream = false;
    return CWSL_OK;
}

int cwsl_rx_enable_timing(cwsl_rx_t* rx, int on) {
    if (!rx) return fail(CWSL_ERR_INVALID, "null receiver");
    rx->timing = on != 0;
    return CWSL_OK;
}

int cwsl_rx_set_stft_guard(cwsl_rx_t* rx, double db_below_band_power) {
    if (!rx || !(db_below_band_power >= 0) || db_below_band_power > 200) return fail(CWSL_ERR_INVALID, "bad guard threshold");
    rx->guard_db = db_below_band_power;
    return CWSL_OK;
}

int cwsl_rx_add_group(cwsl_rx_t* rx, double period_s) {
    if (!rx) return fail(CWSL_ERR_INVALID, "null receiver");
    if (rx->committed) return fail(CWSL_ERR_STATE, "groups must be added before the first push");
    if (!(period_s > 0) || period_s > 86400) return fail(CWSL_ERR_INVALID, "bad period %g", period_s);
    Group g;
    g.period = period_s;
    g.af_size = cwsl_af_size(period_s);
    if (g.af_size <= (size_t)rx->iq_len + 1)
        return fail(CWSL_ERR_INVALID, "period %g s too short for iq_len %u", period_s, rx->iq_len);
    rx->groups.push_back(std::move(g));
    return (int)rx->groups.size() - 1;
}

}  // extern "C"

namespace {

// Replace a committed group's channel set between two slots: the new device state is built first (it takes its own
// references on the shared phase tables), then the old one is released, so tables both sets use are never rebuilt.
int rebuild_group(cwsl_rx* rx, Group& g, std::vector<ChannelHost> channels) {
    if (channels.empty()) return fail(CWSL_ERR_STATE, "a slot group cannot lose its last channel");
    CK(cudaStreamSynchronize(rx->stream));       // nothing may still read the old buffers
    CK(cudaStreamSynchronize(rx->copy_stream));
    rx->d2h_pending = false;
    Group old = g;                               // (shallow: device pointers and cache references move to `old`)
    g.ch = std::move(channels);
    g.phase_keys.clear();
    g.have_anchor_ref = false;
    g.d_tone = nullptr;
    g.d_phase = nullptr;
    g.d_pinc = nullptr;
    g.have_phase_tables = false;
    g.d_sign = g.d_scale = g.d_factor = g.d_maxval = g.d_audio = g.d_seg_scale = nullptr;
    g.d_maxbits = g.d_seg_max = g.d_seg_energy = g.d_n_items = nullptr;
    g.d_sel = nullptr;
    g.d_items = nullptr;
    g.d_counters = nullptr;
    g.d_out = nullptr;
    g.d_chan = nullptr;
    g.d_anchors = nullptr;
    g.max_segs = 0;
    g.committed = false;
    g.have_result = false;
    const int rc = commit_group(rx, g);
    if (rc != CWSL_OK) {                         // keep the old channel set alive
        const std::string msg = g_last_error;
        g = old;
        g_last_error = msg;
        return rc;
    }
    free_group_device(old);
    // ring sizing depends on the group's segment length, which depends on its size: a grown group may need more
    if (rx->d_ring && !rx->bound) {
        uint64_t seg = 0;
        for (const Group& h : rx->groups) seg = std::max<uint64_t>(seg, cwsl::fast_seg_blocks(h.tiles));
        if (rx->own_ring_blocks < 2 * seg + 4 * (uint64_t)rx->sub + 64)
            return fail(CWSL_ERR_STATE, "IQ ring too short for the grown slot group: create the receiver with ring_seconds >= %.2f",
                        (double)(2 * seg + 4 * (uint64_t)rx->sub + 64) * rx->geo.block_size / rx->fs);
    }
    return CWSL_OK;
}

int apply_pending(cwsl_rx* rx, Group& g) {
    if (g.pending_add.empty() && g.pending_remove.empty()) return CWSL_OK;
    std::vector<ChannelHost> next;
    std::set<int> gone(g.pending_remove.begin(), g.pending_remove.end());
    for (size_t c = 0; c < g.ch.size(); ++c)
        if (!gone.count((int)c)) next.push_back(g.ch[c]);
    for (ChannelHost& c : g.pending_add) next.push_back(std::move(c));
    g.pending_add.clear();
    g.pending_remove.clear();
    return rebuild_group(rx, g, std::move(next));
}

}  // namespace

extern "C" {

int cwsl_rx_add_channel(cwsl_rx_t* rx, int group, int32_t demod_freq_hz, int is_usb, float scale) {
    Group* g = get_group(rx, group);
    if (!g) return CWSL_ERR_INVALID;
    if (!(scale > 0.0f) || scale > 1.0f)  // source/CWSL_DIGI.cpp:952-978
        return fail(CWSL_ERR_INVALID, "audio scale factor %g outside (0, 1]", (double)scale);
    ChannelHost ch;
    ch.demod_freq = demod_freq_hz;
    ch.usb = is_usb != 0;
    ch.scale = scale;
    if (!cwsl::nco_tables(rx->geo, demod_freq_hz, is_usb != 0, &ch.nco))
        return fail(CWSL_ERR_INVALID, "Signal outside of band (demod %d Hz at Fs %u)", demod_freq_hz, rx->fs);
    const size_t after = g->ch.size() - g->pending_remove.size() + g->pending_add.size();
    if (after >= 65535)  // channels are a grid dimension of the normalise/quantise launch
        return fail(CWSL_ERR_INVALID, "at most 65535 channels per slot group");
    if (!rx->committed) {
        g->ch.push_back(std::move(ch));
        return (int)g->ch.size() - 1;
    }
    // Running receiver (a decoder instance restarted on this band, source/CWSL_DIGI.cpp:1217-1226): the channel joins
    // at the group's next slot edge; between slots that is now.
    DeviceGuard dg(rx->device);
    if (!dg.ok) return fail(CWSL_ERR_CUDA, "cudaSetDevice(%d) failed", rx->device);
    g->pending_add.push_back(std::move(ch));
    const int index = (int)after;
    if (g->iq_blocks == 0 && g->processed == 0) {
        const int rc = apply_pending(rx, *g);
        if (rc != CWSL_OK) return rc;
    }
    return index;
}

int cwsl_rx_remove_channel(cwsl_rx_t* rx, int group, int channel) {
    Group* g = get_group(rx, group);
    if (!g) return CWSL_ERR_INVALID;
    if (channel < 0 || channel >= (int)g->ch.size()) return fail(CWSL_ERR_INVALID, "bad channel %d", channel);
    if (std::count(g->pending_remove.begin(), g->pending_remove.end(), channel))
        return fail(CWSL_ERR_STATE, "channel %d is already being removed", channel);
    if (g->ch.size() - g->pending_remove.size() + g->pending_add.size() <= 1)
        return fail(CWSL_ERR_STATE, "a slot group cannot lose its last channel");
    if (!rx->committed) {
        g->ch.erase(g->ch.begin() + channel);
        return CWSL_OK;
    }
    DeviceGuard dg(rx->device);
    if (!dg.ok) return fail(CWSL_ERR_CUDA, "cudaSetDevice(%d) failed", rx->device);
    g->pending_remove.push_back(channel);
    if (g->iq_blocks == 0 && g->processed == 0) return apply_pending(rx, *g);
    return CWSL_OK;
}

int cwsl_rx_num_groups(const cwsl_rx_t* rx) { return rx ? (int)rx->groups.size() : CWSL_ERR_INVALID; }

int cwsl_rx_num_channels(const cwsl_rx_t* rx, int group) {
    if (!rx || group < 0 || group >= (int)rx->groups.size()) return CWSL_ERR_INVALID;
    return (int)rx->groups[group].ch.size();
}

size_t cwsl_rx_group_af_size(const cwsl_rx_t* rx, int group) {
    if (!rx || group < 0 || group >= (int)rx->groups.size()) return 0;
    return rx->groups[group].af_size;
}

int cwsl_rx_push_iq(cwsl_rx_t* rx, const float* iq, size_t n_blocks) {
    return push_common(rx, iq, n_blocks, cudaMemcpyHostToDevice);
}

int cwsl_rx_push_iq_device(cwsl_rx_t* rx, const float* d_iq, size_t n_blocks) {
    return push_common(rx, d_iq, n_blocks, cudaMemcpyDeviceToDevice);
}

int cwsl_rx_push_fence(cwsl_rx_t* rx, uint64_t* token) {
    if (!rx || !token) return fail(CWSL_ERR_INVALID, "bad arguments");
    DeviceGuard dg(rx->device);
    if (!dg.ok) return fail(CWSL_ERR_CUDA, "cudaSetDevice(%d) failed", rx->device);
    cudaEvent_t& e = rx->fence_ev[rx->fence_next % cwsl_rx::kFences];
    if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaEventRecord(e, rx->stream));  // (re-recording a slot: its old token is kFences fences older, see wait)
    *token = rx->fence_next++;
    return CWSL_OK;
}

int cwsl_rx_wait_fence(cwsl_rx_t* rx, uint64_t token) {
    if (!rx) return fail(CWSL_ERR_INVALID, "null receiver");
    if (token == 0 || token >= rx->fence_next) return fail(CWSL_ERR_INVALID, "unknown fence %llu", (unsigned long long)token);
    DeviceGuard dg(rx->device);
    // A slot that has been re-recorded since holds a LATER fence of the same stream: waiting for that one is
    // sufficient (stream order), merely longer than necessary.
    CK(cudaEventSynchronize(rx->fence_ev[token % cwsl_rx::kFences]));
    return CWSL_OK;
}

int cwsl_rx_bind_device_iq(cwsl_rx_t* rx, const float* d_iq, size_t n_blocks) {
    if (!rx || !d_iq || n_blocks == 0) return fail(CWSL_ERR_INVALID, "bad arguments");
    if ((reinterpret_cast<uintptr_t>(d_iq) & 15u) != 0) return fail(CWSL_ERR_INVALID, "IQ buffer must be 16-byte aligned");
    DeviceGuard dg(rx->device);
    if (!dg.ok) return fail(CWSL_ERR_CUDA, "cudaSetDevice(%d) failed", rx->device);
    int rc = commit(rx);
    if (rc != CWSL_OK) return rc;
    rx->bound = true;
    const uint64_t blocks = (uint64_t)n_blocks * rx->sub;
    if (blocks > 0xFFFFFFF0ull) return fail(CWSL_ERR_INVALID, "slot too long");
    rx->ring_ptr = reinterpret_cast<const float2*>(d_iq);
    rx->ring_blocks = (uint32_t)blocks;
    rx->abs_written = blocks;
    for (Group& g : rx->groups) {
        g.slot_start = 0;
        g.processed = 0;
        g.iq_blocks = n_blocks;
    }
    return CWSL_OK;
}

int cwsl_rx_process(cwsl_rx_t* rx, int group) {
    if (!rx) return fail(CWSL_ERR_INVALID, "null receiver");
    DeviceGuard dg(rx->device);
    if (!dg.ok) return fail(CWSL_ERR_CUDA, "cudaSetDevice(%d) failed", rx->device);
    int rc = commit(rx);
    if (rc != CWSL_OK) return rc;
    if (group >= 0) {
        Group* g = get_group(rx, group);
        return g ? process_group(rx, *g, false) : CWSL_ERR_INVALID;
    }
    for (Group& g : rx->groups)
        if ((rc = process_group(rx, g, false)) != CWSL_OK) return rc;
    return CWSL_OK;
}

}  // extern "C"

namespace {

// Kernels and copies of a slot edge; the caller resets the slot whatever this returns.
int end_slot_work(cwsl_rx* rx, Group* g, int16_t* out_i16, bool packed) {
    if (const char* e = std::getenv("CWSL_TEST_FAIL_END_SLOT"); e && e[0] == '1')  // fault injection for the tests
        return fail(CWSL_ERR_CUDA, "injected failure (CWSL_TEST_FAIL_END_SLOT)");
    int rc = process_group(rx, *g, true);
    if (rc != CWSL_OK) return rc;
    const uint32_t C = (uint32_t)g->ch.size();
    cwsl::QuantLaunch q;
    q.audio = g->d_audio;
    q.af_stride = g->af_stride;
    q.n_channels = C;
    q.write_index = (uint32_t)g->processed;
    q.af_size = (uint32_t)g->af_size;
    q.cover = (uint32_t)std::max<size_t>(g->out_dirty, (size_t)g->processed);
    // PACKED hand-off: rows of write_index samples back to back, no zero tail -- the int16 result is then ONE contiguous
    // range, which crosses PCIe as a 1-D copy (55.6 GB/s on B200 / Gen5 x16 against 51.9 for the strided 2-D copy)
    q.out_pitch = packed ? (uint32_t)g->processed : (uint32_t)g->af_size;
    q.maxbits = g->d_maxbits;
    q.scale = g->d_scale;
    q.out = g->d_out;
    q.factor_out = g->d_factor;
    q.max_out = g->d_maxval;
    // Everything after the demodulation -- the HBM-bound normalise/quantise pass, the max reset and the copy to
    // the host -- runs on the receiver's private post stream behind an event, so it overlaps the FMA-bound
    // demodulation of whatever is queued next on `stream` (other receivers sharing it, or this receiver's other
    // groups). Work of successive slots is ordered on the post stream itself; the next demodulation of THIS
    // receiver waits for ev_d2h_done (process_group), because it rewrites the float audio and the max.
    CK(cudaEventRecord(rx->ev_out_ready, rx->stream));
    CK(cudaStreamWaitEvent(rx->copy_stream, rx->ev_out_ready, 0));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (rx->timing) {
        e0 = get_event(rx);
        e1 = get_event(rx);
        CK(cudaEventRecord(e0, rx->copy_stream));
    }
    CK(cwsl::launch_quantise(q, rx->copy_stream));
    // (after a packed slot nothing is known about the [n][af_size] view of the buffer: the next unpacked slot rewrites it all)
    g->out_dirty = packed ? g->af_size : (size_t)g->processed;
    g->out_pitch = q.out_pitch;
    // next slot starts from max|x| = 0; the guard's counters roll over to "last finished slot"
    CK(cwsl::launch_clear_u32(g->d_maxbits, C, g->d_counters, rx->copy_stream));
    if (rx->timing) {
        CK(cudaEventRecord(e1, rx->copy_stream));
        rx->ev_quant.emplace_back(e0, e1);
    }
    if (out_i16 && packed) {
        {   // a packed result inside a managed region overwrites whatever [n][af_size] views were tracked there
            std::lock_guard<std::mutex> lk(g_host_mu);
            const uintptr_t a = reinterpret_cast<uintptr_t>(out_i16);
            auto it = g_host_regions.upper_bound(a);
            if (it != g_host_regions.begin()) {
                --it;
                if (a >= it->first && a < it->first + it->second.bytes)
                    for (auto& kv : it->second.outs) kv.second.dirty_cols = kv.second.af_size;
            }
        }
        if (g->processed > 0)
            CK(cudaMemcpyAsync(out_i16, g->d_out, (size_t)C * g->processed * sizeof(int16_t), cudaMemcpyDeviceToHost,
                               rx->copy_stream));
    } else if (out_i16) {
        size_t cols = g->af_size;  // default: the whole buffer, zero tail included
        {
            std::lock_guard<std::mutex> lk(g_host_mu);
            const uintptr_t a = reinterpret_cast<uintptr_t>(out_i16);
            auto it = g_host_regions.upper_bound(a);
            if (it != g_host_regions.begin()) {
                --it;
                const size_t need = (size_t)C * g->af_size * sizeof(int16_t);
                if (a >= it->first && a + need <= it->first + it->second.bytes) {
                    OutState& st = it->second.outs[a];
                    if (st.af_size != g->af_size || st.rows != C) {
                        // first use of this destination (or a different shape): drop overlapping records, the
                        // region was zeroed at allocation so only what they wrote can be dirty -> full copy once
                        const bool fresh = st.af_size == 0 && it->second.outs.size() == 1;
                        if (!fresh)  // another destination in this region may overlap: nothing is known any more
                            for (auto& kv : it->second.outs) kv.second.dirty_cols = kv.second.af_size;
                        st.af_size = g->af_size;
                        st.rows = C;
                        st.dirty_cols = fresh ? 0 : g->af_size;
                    }
                    cols = std::max<size_t>(st.dirty_cols, (size_t)g->processed);
                    st.dirty_cols = (size_t)g->processed;
                }
            }
        }
        if (cols >= g->af_size) {
            CK(cudaMemcpyAsync(out_i16, g->d_out, (size_t)C * g->af_size * sizeof(int16_t), cudaMemcpyDeviceToHost,
                               rx->copy_stream));
        } else if (cols > 0) {
            CK(cudaMemcpy2DAsync(out_i16, g->af_size * sizeof(int16_t), g->d_out, g->af_size * sizeof(int16_t),
                                 cols * sizeof(int16_t), C, cudaMemcpyDeviceToHost, rx->copy_stream));
        }
    }
    CK(cudaEventRecord(rx->ev_d2h_done, rx->copy_stream));
    rx->d2h_pending = true;
    return CWSL_OK;
}

}  // namespace

extern "C" {

static int end_slot_common(cwsl_rx_t* rx, int group, int16_t* out_i16, size_t* write_index, bool packed);

int cwsl_rx_end_slot(cwsl_rx_t* rx, int group, int16_t* out_i16, size_t* write_index) {
    return end_slot_common(rx, group, out_i16, write_index, false);
}

int cwsl_rx_end_slot_packed(cwsl_rx_t* rx, int group, int16_t* out_i16, size_t* write_index) {
    return end_slot_common(rx, group, out_i16, write_index, true);
}

static int end_slot_common(cwsl_rx_t* rx, int group, int16_t* out_i16, size_t* write_index, bool packed) {
    Group* g = get_group(rx, group);
    if (!g) return CWSL_ERR_INVALID;
    DeviceGuard dg(rx->device);
    if (!dg.ok) return fail(CWSL_ERR_CUDA, "cudaSetDevice(%d) failed", rx->device);
    int rc = commit(rx);
    if (rc != CWSL_OK) return rc;
    rc = end_slot_work(rx, g, out_i16, packed);
    if (rc == CWSL_OK) {
        if (write_index) *write_index = (size_t)g->processed;
        g->last_write_index = (size_t)g->processed;
        g->have_result = true;
    } else {
        // A failed slot is LOST, not stuck: the reference logs the error and carries on with the next slot
        // (source/Receiver.hpp:222-229, source/Instance.cpp:268-271). Whatever part of it reached the max is dropped.
        const std::string msg = g_last_error;
        if (write_index) *write_index = 0;
        g->have_result = false;
        if (g->d_maxbits) cudaMemsetAsync(g->d_maxbits, 0, g->ch.size() * sizeof(unsigned), rx->stream);
        cudaGetLastError();
        g_last_error = msg;
    }
    // slot reset: fresh SSBD per slot (source/Instance.cpp:251) -- on the error path as well
    g->slot_start = rx->abs_written;
    g->processed = 0;
    g->iq_blocks = 0;
    if (rc == CWSL_OK && (!g->pending_add.empty() || !g->pending_remove.empty())) {
        // the finished slot's int16 result must stay readable: it is copied out before the buffers are replaced
        if (!out_i16) return fail(CWSL_ERR_STATE, "channel-set change pending: this slot's result must be taken to the host");
        rc = apply_pending(rx, *g);
    }
    return rc;
}

const int16_t* cwsl_rx_device_audio(const cwsl_rx_t* rx, int group) {
    if (!rx || group < 0 || group >= (int)rx->groups.size()) return nullptr;
    return rx->groups[group].d_out;
}

int cwsl_rx_copy_device_audio(cwsl_rx_t* rx, int group, int channel, int16_t* d_dst) {
    Group* g = get_group(rx, group);
    if (!g || !d_dst) return fail(CWSL_ERR_INVALID, "bad arguments");
    if (!g->have_result) return fail(CWSL_ERR_STATE, "no finished slot available");
    if (channel < 0 || channel >= (int)g->ch.size()) return fail(CWSL_ERR_INVALID, "bad channel %d", channel);
    DeviceGuard dg(rx->device);
    if (g->out_pitch && g->out_pitch != g->af_size) {  // packed slot: the row holds write_index samples, the tail is zero
        const size_t wi = g->out_pitch;
        if (wi) CK(cudaMemcpyAsync(d_dst, g->d_out + (size_t)channel * wi, wi * sizeof(int16_t), cudaMemcpyDeviceToDevice, rx->copy_stream));
        CK(cudaMemsetAsync(d_dst + wi, 0, (g->af_size - wi) * sizeof(int16_t), rx->copy_stream));
        return CWSL_OK;
    }
    CK(cudaMemcpyAsync(d_dst, g->d_out + (size_t)channel * g->af_size, g->af_size * sizeof(int16_t),
                       cudaMemcpyDeviceToDevice, rx->copy_stream));  // ordered behind the slot's quantise
    return CWSL_OK;
}

int cwsl_rx_read_float_audio(cwsl_rx_t* rx, int group, int channel, float* out) {
    Group* g = get_group(rx, group);
    if (!g || !out) return CWSL_ERR_INVALID;
    if (!g->have_result || g->processed != 0) return fail(CWSL_ERR_STATE, "no finished slot available");
    if (channel < 0 || channel >= (int)g->ch.size()) return fail(CWSL_ERR_INVALID, "bad channel %d", channel);
    DeviceGuard dg(rx->device);
    CK(cudaStreamSynchronize(rx->stream));
    CK(cudaStreamSynchronize(rx->copy_stream));
    std::memset(out, 0, g->af_size * sizeof(float));
    CK(cudaMemcpy(out, g->d_audio + (size_t)channel * g->af_stride, g->last_write_index * sizeof(float),
                  cudaMemcpyDeviceToHost));
    return CWSL_OK;
}

int cwsl_rx_channel_stats(cwsl_rx_t* rx, int group, int channel, float* max_val, float* factor) {
    Group* g = get_group(rx, group);
    if (!g) return CWSL_ERR_INVALID;
    if (!g->have_result || g->processed != 0) return fail(CWSL_ERR_STATE, "no finished slot available");
    if (channel < 0 || channel >= (int)g->ch.size()) return fail(CWSL_ERR_INVALID, "bad channel %d", channel);
    DeviceGuard dg(rx->device);
    CK(cudaStreamSynchronize(rx->stream));
    CK(cudaStreamSynchronize(rx->copy_stream));
    if (max_val) CK(cudaMemcpy(max_val, g->d_maxval + channel, sizeof(float), cudaMemcpyDeviceToHost));
    if (factor) CK(cudaMemcpy(factor, g->d_factor + channel, sizeof(float), cudaMemcpyDeviceToHost));
    return CWSL_OK;
}

int cwsl_rx_synchronize(cwsl_rx_t* rx) {
    if (!rx) return fail(CWSL_ERR_INVALID, "null receiver");
    DeviceGuard dg(rx->device);
    CK(cudaStreamSynchronize(rx->stream));
    CK(cudaStreamSynchronize(rx->copy_stream));
    rx->d2h_pending = false;
    return CWSL_OK;
}

int cwsl_rx_wait_output(cwsl_rx_t* rx) {
    if (!rx) return fail(CWSL_ERR_INVALID, "null receiver");
    DeviceGuard dg(rx->device);
    if (rx->d2h_pending) {
        CK(cudaEventSynchronize(rx->ev_d2h_done));
        rx->d2h_pending = false;
    }
    return CWSL_OK;
}

int cwsl_rx_join_output(cwsl_rx_t* rx) {
    if (!rx) return fail(CWSL_ERR_INVALID, "null receiver");
    DeviceGuard dg(rx->device);
    if (rx->d2h_pending) CK(cudaStreamWaitEvent(rx->stream, rx->ev_d2h_done, 0));
    return CWSL_OK;
}

void* cwsl_rx_stream(cwsl_rx_t* rx) { return rx ? (void*)rx->stream : nullptr; }

int cwsl_rx_kernel_times_ex(cwsl_rx_t* rx, float ms[5], int launches[2]) {
    if (!rx) return fail(CWSL_ERR_INVALID, "null receiver");
    DeviceGuard dg(rx->device);
    CK(cudaStreamSynchronize(rx->stream));
    CK(cudaStreamSynchronize(rx->copy_stream));
    float dsum = 0, qsum = 0, pre = 0, main_ms = 0, post = 0;
    for (auto& ev : rx->ev_demod) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, ev.e0, ev.e1));
        dsum += t;
        if (ev.e_pre && ev.e_main) {  // STFT launch: guard prologue | channelizer | guard selection + direct-form redo
            CK(cudaEventElapsedTime(&t, ev.e0, ev.e_pre));
            pre += t;
            CK(cudaEventElapsedTime(&t, ev.e_pre, ev.e_main));
            main_ms += t;
            CK(cudaEventElapsedTime(&t, ev.e_main, ev.e1));
            post += t;
        } else {
            CK(cudaEventElapsedTime(&t, ev.e0, ev.e1));
            main_ms += t;
        }
        for (cudaEvent_t e : {ev.e0, ev.e_pre, ev.e_main, ev.e1})
            if (e) rx->ev_pool.push_back(e);
    }
    for (auto& pr : rx->ev_quant) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, pr.first, pr.second));
        qsum += t;
        rx->ev_pool.push_back(pr.first);
        rx->ev_pool.push_back(pr.second);
    }
    if (ms) {
        ms[0] = dsum;
        ms[1] = qsum;
        ms[2] = main_ms;
        ms[3] = pre;
        ms[4] = post;
    }
    if (launches) {
        launches[0] = (int)rx->ev_demod.size();
        launches[1] = (int)rx->ev_quant.size();
    }
    rx->ev_demod.clear();
    rx->ev_quant.clear();
    return CWSL_OK;
}

int cwsl_rx_kernel_times(cwsl_rx_t* rx, float* demod_ms, float* quant_ms, int* demod_launches, int* quant_launches) {
    float ms[5] = {0, 0, 0, 0, 0};
    int n[2] = {0, 0};
    const int rc = cwsl_rx_kernel_times_ex(rx, ms, n);
    if (rc != CWSL_OK) return rc;
    if (demod_ms) *demod_ms = ms[0];
    if (quant_ms) *quant_ms = ms[1];
    if (demod_launches) *demod_launches = n[0];
    if (quant_launches) *quant_launches = n[1];
    return CWSL_OK;
}

int cwsl_rx_guard_stats(cwsl_rx_t* rx, int group, uint64_t* decided, uint64_t* redone) {
    Group* g = get_group(rx, group);
    if (!g) return CWSL_ERR_INVALID;
    if (!g->have_result) return fail(CWSL_ERR_STATE, "no finished slot available");
    DeviceGuard dg(rx->device);
    CK(cudaStreamSynchronize(rx->stream));
    CK(cudaStreamSynchronize(rx->copy_stream));
    unsigned long long v[4] = {0, 0, 0, 0};
    if (g->d_counters) CK(cudaMemcpy(v, g->d_counters, sizeof(v), cudaMemcpyDeviceToHost));
    if (decided) *decided = v[2];
    if (redone) *redone = v[3];
    return CWSL_OK;
}

void* cwsl_host_alloc(size_t bytes) {
    if (bytes == 0) {
        fail(CWSL_ERR_INVALID, "zero-size allocation");
        return nullptr;
    }
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        fail(CWSL_ERR_NOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    std::memset(p, 0, bytes);
    std::lock_guard<std::mutex> lk(g_host_mu);
    g_host_regions[reinterpret_cast<uintptr_t>(p)].bytes = bytes;
    return p;
}

void cwsl_host_free(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        g_host_regions.erase(reinterpret_cast<uintptr_t>(p));
    }
    cudaFreeHost(p);
}

int cwsl_measure_fp32_peak(int device, float* ffma_tflops, float* ffma2_tflops) {
    int n = cwsl_device_count();
    if (n <= 0) return fail(CWSL_ERR_CUDA, "no CUDA device available");
    if (device < 0 || device >= n) return fail(CWSL_ERR_INVALID, "device %d out of range", device);
    DeviceGuard dg(device);
    CK(cwsl::measure_fp32_peak(ffma_tflops, ffma2_tflops));
    return CWSL_OK;
}

}  // extern "C"
