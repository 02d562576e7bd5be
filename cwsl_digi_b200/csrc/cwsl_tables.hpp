// Host-side table builder of the B200 front-end.
//
// The demodulator kernels consume exactly the tables the reference's SSBD<float> object holds,
// so they are produced here on the host by the same libm calls and the same expression shapes
// (rounding points) as the reference:
//   low-pass prototype   source/LowPass.hpp:16-35   (double math, one rounding to float per tap)
//   DC normalisation     source/SSBD.hpp:66-68      (sequential float sum, float divide)
//   NCO tone / phase_inc source/SSBD.hpp:110-114    (float phase_delta, glibc cexpf)
// tests/test_capi_cpu.py (test_host_tables_bit_identical_to_reference, test_tables_property) pins them bit-for-bit against oracle/_ref (the reference's own headers).
#pragma once

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace cwsl {

constexpr uint32_t kWaveSR = 12000;   // Wave_SR, source/CWSL_DIGI.hpp:51
constexpr uint32_t kSSBBW = 6000;     // SSB_BW,  source/CWSL_DIGI.hpp:52
constexpr uint32_t kLatencyLog2 = 3;  // SSBD ctor default, source/SSBD.hpp:49
constexpr uint32_t kNumWS = 32;       // FiltOrder/BlockSize = 4*latency for every legal Fs
constexpr double kPi = 3.14159265358979323846;  // source/LowPass.hpp:13

struct SsbdGeometry {
    uint32_t fs = 0;
    uint32_t filt_order = 0;  // latency*2*Fs/B   source/SSBD.hpp:62
    uint32_t block_size = 0;  // Fs/B/2           source/SSBD.hpp:71
    uint32_t num_ws = 0;      // FiltOrder/BlockSize
    uint32_t in_size = 0;     // 2*Fs/B           source/SSBD.hpp:144
    uint32_t dec_ratio = 0;   // Fs/Wave_SR       source/Instance.cpp:192
};

// false where the SSBD ctor throws (source/SSBD.hpp:54-59)
inline bool ssbd_geometry(uint32_t fs, SsbdGeometry* g) {
    const size_t Fs = fs, B = kSSBBW;
    if ((Fs / B / 2) * 2 * B != Fs || Fs < 4 * B) return false;
    const size_t latency = size_t(1) << kLatencyLog2;
    g->fs = fs;
    g->filt_order = uint32_t(latency * 2 * Fs / B);
    g->block_size = uint32_t(Fs / B / 2);
    g->num_ws = g->filt_order / g->block_size;
    g->in_size = uint32_t(2 * Fs / B);
    g->dec_ratio = fs / kWaveSR;
    return true;
}

// Hamming-windowed sinc prototype, normalised to unit DC gain.
inline std::vector<float> lowpass_taps(const SsbdGeometry& g) {
    const size_t order = g.filt_order;
    const double bandwidth = kSSBBW / (double)g.fs;
    std::vector<float> f(order);
    f[0] = static_cast<float>(0.0);
    f[order / 2] = static_cast<float>(1.0);
    const double x0 = -1.0 * order / 2;
    for (size_t n = 1; n < order / 2; ++n) {
        const double xPi = (x0 + n) * kPi * bandwidth;
        const double y = sin(xPi) / xPi * (0.54 - 0.46 * cos(2.0 * kPi * n / (double)order));
        f[n] = static_cast<float>(y);
        f[order - n] = static_cast<float>(y);
    }
    float sum = 0.0;
    for (size_t n = 0; n < order; ++n) sum += f[n];
    for (size_t n = 0; n < order; ++n) f[n] /= sum;
    return f;
}

struct NcoTables {
    std::vector<std::complex<float>> tone;  // [block_size]
    std::complex<float> phase_inc;
    float sign = 1.0f;
    float phase_delta = 0.0f;               // radians per input sample, source/SSBD.hpp:111
};

// false where SSBD::Tune throws (source/SSBD.hpp:100-103)
inline bool nco_tables(const SsbdGeometry& g, int32_t demod_freq_hz, bool is_usb, NcoTables* t) {
    const size_t Fs = g.fs, B = kSSBBW;
    const double F = static_cast<float>(demod_freq_hz);  // source/Instance.cpp:187 passes a float
    if (fabs(F) > Fs / 2) return false;
    if (fabs(F + B * (is_usb ? 1.0 : -1.0)) > Fs / 2) return false;
    const float sign = static_cast<float>(is_usb ? 1.0 : -1.0);
    const float phase_delta = static_cast<float>(-2.0 * kPi * (F + sign * B / 2.0) / static_cast<double>(Fs));
    t->sign = sign;
    t->phase_delta = phase_delta;
    t->tone.resize(g.block_size);
    for (size_t n = 0; n < g.block_size; ++n)
        t->tone[n] = std::exp(std::complex<float>(0.0, phase_delta * n));
    t->phase_inc = std::exp(std::complex<float>(0.0, phase_delta * g.block_size));
    return true;
}

// ---- STFT channelizer tables (cwsl_chan.cu). Not reference arithmetic: a type-2 non-uniform FFT of the
// reference's window (512 taps on a 1024-point grid) with a Kaiser-Bessel interpolation kernel, all in double.
inline uint32_t chan_grid(const SsbdGeometry& g) { return 2 * g.filt_order; }  // 1024 / 512 / 256 bins

struct ChanKernel {
    int w;
    double beta, i0beta;
    explicit ChanKernel(int taps) : w(taps) {
        const double sigma = 2.0;  // grid / window length
        beta = kPi * std::sqrt((w / sigma) * (w / sigma) * (sigma - 0.5) * (sigma - 0.5) - 0.8);
        i0beta = std::cyl_bessel_i(0.0, beta);
    }
    double psi(double t) const {  // interpolation kernel, support |t| <= w/2 bins
        const double a = 1.0 - (2.0 * t / w) * (2.0 * t / w);
        return a < 0.0 ? 0.0 : std::cyl_bessel_i(0.0, beta * std::sqrt(a)) / i0beta;
    }
    double psihat(double s) const {  // its Fourier transform at s cycles per bin (|pi w s| < beta here)
        const double z2 = beta * beta - (kPi * w * s) * (kPi * w * s);
        const double z = std::sqrt(z2);
        return w * std::sinh(z) / z / i0beta;
    }
};

// window[j] = h[j] / psihat((j - L/2)/N): the low-pass taps pre-compensated for the interpolation kernel
inline std::vector<float> chan_window(const SsbdGeometry& g, int width) {
    const std::vector<float> h = lowpass_taps(g);
    const ChanKernel k(width);
    std::vector<float> w(h.size());
    for (size_t j = 0; j < h.size(); ++j)
        w[j] = static_cast<float>((double)h[j] / k.psihat(((double)j - h.size() / 2.0) / chan_grid(g)));
    return w;
}

// twiddle[q1*32 + j2] = W_N^(j2*q1) * i^q1, q1 < N/32 (inter-pass twiddle of the (N/32) x 32 FFT and the rotation
// that moves the phase reference of the spectrum to the window centre, e^{+2 pi i q (L/2)/N} = i^q)
inline std::vector<std::complex<float>> chan_twiddles(const SsbdGeometry& g) {
    const int n = (int)chan_grid(g), a = n / 32;
    std::vector<std::complex<float>> t((size_t)a * 32);
    for (int q1 = 0; q1 < a; ++q1)
        for (int j2 = 0; j2 < 32; ++j2) {
            const double ang = -2.0 * kPi * ((j2 * q1) % n) / n + kPi / 2.0 * (q1 & 3);
            t[(size_t)q1 * 32 + j2] = std::complex<float>((float)std::cos(ang), (float)std::sin(ang));
        }
    return t;
}

struct ChanChannel {
    int q0 = 0;                  // first bin of the stencil (may be negative: bins are taken mod N)
    std::vector<float> wgt;      // [taps]
    std::complex<float> rot;     // e^{-i 240 w}
};

// Per-channel constants: the channel's effective NCO frequency is the per-block angle of the reference's own
// float phase_inc (unwrapped with phase_delta) divided by the block size.
// The stencil is `taps` bins starting at an even bin and covers the support of the width-`width` kernel
// (taps >= width + 1), so the kernel can read it as aligned bin pairs.
inline ChanChannel chan_channel(const SsbdGeometry& g, const NcoTables& t, int width, int taps) {
    const ChanKernel k(width);
    double theta = std::atan2((double)t.phase_inc.imag(), (double)t.phase_inc.real());
    const double nominal = (double)t.phase_delta * g.block_size;
    theta += 2.0 * kPi * std::round((nominal - theta) / (2.0 * kPi));
    const double omega = theta / g.block_size;
    const double n = chan_grid(g);
    double nu = std::fmod(-omega * n / (2.0 * kPi), n);
    if (nu < 0) nu += n;
    ChanChannel c;
    c.q0 = (int)std::ceil(nu - width / 2.0);
    if (c.q0 & 1) c.q0 -= 1;
    c.wgt.resize(taps);
    for (int i = 0; i < taps; ++i) c.wgt[i] = (float)k.psi(nu - (c.q0 + i));
    const double a = -omega * (double)(g.filt_order - g.block_size - g.filt_order / 2);
    c.rot = std::complex<float>((float)std::cos(a), (float)std::sin(a));
    return c;
}

// ---- interpolation work items of the channelizer kernel (cwsl_chan.cu) ----------------------------------------
// The kernel reads a channel's frequency off the FFT grid with the bins of the kernel's support [a, a+6],
// a = ceil(nu - width/2). Channels that are neighbours on the grid share most of those bins, so the kernel works on
// ITEMS of up to 4 channels ("members") that read ONE 12-bin window (even first bin e) from shared memory:
// member j uses the 9 bins e + shift[j] ... e + shift[j] + 8 with shift = {0, 0, 1, 2} (compile-time register
// indices), weights outside the kernel's support are exact zeros. A member fits slot j iff shift[j] <= a - e <=
// shift[j] + 2. Channels are taken in ascending grid position and packed greedily; any channel set is legal (a
// channel that fits no open slot starts a new item), a regular grid of 130 ... 190 Hz spacing fills all four slots.
// The result of a channel does not depend on how it was grouped: its non-zero weights meet the same bins in the same
// (ascending) order, and a fused multiply-add with an exact-zero weight is the identity.
constexpr int kItemMembers = 4, kItemTaps = 9, kItemBins = 12;
constexpr int kItemShift[kItemMembers] = {0, 0, 1, 2};
struct ChanItemHost {
    int e = 0;                              // first bin of the window (even; may be negative: bins are taken mod N)
    int n = 0;                              // members in use (slots may stay empty in between)
    int ch[kItemMembers] = {-1, -1, -1, -1};  // channel index in the slot group, -1: empty slot
    float w[kItemMembers][kItemTaps] = {};
    std::complex<float> rot[kItemMembers];
};

struct ChanPos {  // grid position of one channel: nu in [0, N) bins, support start a, rotation e^{-i 240 w}
    double nu = 0;
    int a = 0;
    std::complex<float> rot;
};
inline ChanPos chan_position(const SsbdGeometry& g, const NcoTables& t, int width) {
    double theta = std::atan2((double)t.phase_inc.imag(), (double)t.phase_inc.real());
    const double nominal = (double)t.phase_delta * g.block_size;
    theta += 2.0 * kPi * std::round((nominal - theta) / (2.0 * kPi));
    const double omega = theta / g.block_size;
    const double n = chan_grid(g);
    ChanPos p;
    p.nu = std::fmod(-omega * n / (2.0 * kPi), n);
    if (p.nu < 0) p.nu += n;
    p.a = (int)std::ceil(p.nu - width / 2.0);
    const double ang = -omega * (double)(g.filt_order - g.block_size - g.filt_order / 2);
    p.rot = std::complex<float>((float)std::cos(ang), (float)std::sin(ang));
    return p;
}

inline std::vector<ChanItemHost> chan_items(const SsbdGeometry& g, const std::vector<NcoTables>& nco, int width) {
    const ChanKernel k(width);
    std::vector<ChanPos> pos(nco.size());
    std::vector<int> order(nco.size());
    for (size_t c = 0; c < nco.size(); ++c) {
        pos[c] = chan_position(g, nco[c], width);
        order[c] = (int)c;
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return pos[x].nu < pos[y].nu; });
    std::vector<ChanItemHost> items;
    int next_slot = kItemMembers;  // no item open
    for (int c : order) {
        const ChanPos& p = pos[c];
        int slot = -1;
        if (!items.empty())
            for (int j = next_slot; j < kItemMembers; ++j) {
                const int d = p.a - items.back().e;
                if (d >= kItemShift[j] && d <= kItemShift[j] + 2) {
                    slot = j;
                    break;
                }
            }
        if (slot < 0) {
            ChanItemHost it;
            it.e = p.a - (p.a & 1);  // (two's complement: also the next lower even number for negative a)
            items.push_back(it);
            slot = 0;
        }
        ChanItemHost& it = items.back();
        it.ch[slot] = c;
        it.rot[slot] = p.rot;
        for (int i = 0; i < kItemTaps; ++i) it.w[slot][i] = (float)k.psi(p.nu - (it.e + kItemShift[slot] + i));
        it.n = slot + 1;
        next_slot = slot + 1;
    }
    return items;
}

// Thread placement of the items of each launch (<= per_launch items, one per interpolation thread). The eight lanes of
// a quarter-warp read 16 bytes each at their window start + 16 q; that is conflict-free iff their starts fall into
// eight different 16-byte bank groups, i.e. (e/2) mod 8 are all different (the same shift q for every lane keeps them
// different). Items are dealt out per launch, quarter-warp by quarter-warp: each of its eight places goes to the
// residue class that is furthest behind an even spread over the quarter-warps still to fill, so a class with more than
// its share costs single 2-way conflicts spread over the launch instead of one fully serialised quarter-warp at the end.
inline std::vector<ChanItemHost> chan_item_order(const std::vector<ChanItemHost>& items, int grid_bins, size_t per_launch) {
    std::vector<ChanItemHost> out;
    out.reserve(items.size());
    for (size_t i0 = 0; i0 < items.size(); i0 += per_launch) {
        const size_t n = std::min(per_launch, items.size() - i0);
        std::vector<std::vector<size_t>> cls(8);
        for (size_t i = i0 + n; i-- > i0;) {  // (reversed, so that pop_back hands them out in ascending order)
            const int e = ((items[i].e % grid_bins) + grid_bins) % grid_bins;
            cls[(e / 2) % 8].push_back(i);
        }
        const size_t quarters = (n + 7) / 8;
        for (size_t q = 0; q < quarters; ++q) {
            const long left = (long)(quarters - q);
            int taken[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const size_t places = std::min<size_t>(8, n - 8 * q);
            for (size_t k = 0; k < places; ++k) {
                int best = -1;
                long best_pri = 0;
                for (int r = 0; r < 8; ++r) {
                    if (cls[r].empty()) continue;
                    const long pri = (long)cls[r].size() + taken[r] - (long)taken[r] * left;  // = count at quarter start - taken * left
                    if (best < 0 || pri > best_pri) {
                        best = r;
                        best_pri = pri;
                    }
                }
                out.push_back(items[cls[best].back()]);
                cls[best].pop_back();
                ++taken[best];
            }
        }
    }
    return out;
}

inline size_t af_size(double period_s) {  // source/Instance.cpp:149
    return static_cast<size_t>(static_cast<double>(kWaveSR) * static_cast<double>(period_s + 5));
}

// source/Instance.cpp:268-276 applied to a run of n_iq_blocks blocks starting from an empty buffer
inline size_t accepted_blocks(size_t n_iq_blocks, uint32_t iq_len, uint32_t dec_ratio, size_t afsize,
                              size_t write_index = 0) {
    size_t acc = 0;
    for (size_t i = 0; i < n_iq_blocks; ++i) {
        if (write_index + iq_len > afsize - 1) break;  // every later block is dropped as well
        write_index += iq_len / dec_ratio;
        ++acc;
    }
    return acc;
}

}  // namespace cwsl
