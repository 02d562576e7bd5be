"""CPU check of the STFT channelizer's host-side tables (cwsl_stft_tables / cwsl_stft_channel, no device needed):
the kernel's arithmetic restated with numpy -- windowed 1024-point FFT per hop (32x32 decomposition with the
library's own twiddle table), 8-bin interpolation, phase table from the oracle -- must reproduce the oracle's audio
to the FAST-mode bars. Pins the math of cwsl_chan.cu on CPU; the kernel itself is checked by the -m gpu tests."""
import numpy as np
import pytest

from cwsl_digi_b200 import synth
from oracle.oracle import af_size

FS, IQ_LEN = 192000, 2048


def geo(fs):
    hop = fs // 12000
    return 2 * 32 * hop, 32 * hop, hop          # grid N, window L, hop (= SSBD block size)


def fft_ax32(u, tw, n):
    """X[q1 + A q2] of the zero-padded window, the way the kernel computes it (an A-point and a 32-point pass with the
    inter-pass twiddles, A = N/32); includes the i^q centring rotation that the twiddle table carries."""
    nb, a = u.shape[0], n // 32
    up = np.zeros((nb, n), np.complex64)
    up[:, :u.shape[1]] = u
    x = up.reshape(nb, a, 32)                        # [j1, j2]
    p1 = np.fft.fft(x, axis=1).astype(np.complex64)  # over j1 -> [q1, j2]
    p1 = p1 * tw[None, :, :]                         # W_N^(j2 q1) * i^q1
    p2 = np.fft.fft(p1, axis=2).astype(np.complex64)  # over j2 -> [q1, q2]
    return np.transpose(p2, (0, 2, 1)).reshape(nb, n)  # bin q1 + A q2


def stft_slot(cw, port, iq, freq, usb=True, fs=FS):
    N, L, HOP = geo(fs)
    x = (iq[0::2] + 1j * iq[1::2]).astype(np.complex64)
    nb = x.size // HOP
    t = cw.stft_tables(fs)
    assert t["grid"] == N and t["window"].size == L
    xpad = np.concatenate([np.zeros(L - HOP, np.complex64), x])
    frames = np.lib.stride_tricks.sliding_window_view(xpad, L)[::HOP][:nb]
    spec = fft_ax32(frames * t["window"][None, :], t["twiddle"], N)
    c = cw.stft_channel(fs, freq, usb)
    acc = (spec[:, (c["q0"] + np.arange(8)) % N] * c["wgt"][None, :]).sum(axis=1)
    tb = port.tables(fs, freq, is_usb=usb)
    ph = port.phase_table(tb["phase_inc"], nb)
    y = acc * ((ph[:, 0] + 1j * ph[:, 1]) * c["rot"])
    sign = 1.0 if usb else -1.0
    b = np.arange(nb) & 3
    return np.where(b == 0, y.real, np.where(b == 1, -y.imag * sign, np.where(b == 2, -y.real, y.imag * sign))).astype(np.float32)


def resid_db(got, want):
    want = want.astype(np.float64)
    err = got.astype(np.float64) - want
    return 20 * np.log10(max(np.sqrt(np.mean(err ** 2)), 1e-300) / np.sqrt(np.mean(want ** 2)))


@pytest.mark.parametrize("fs", [192000, 96000, 48000])
def test_twiddles_are_the_ax32_factors(cw, fs):
    N, L, _ = geo(fs)
    tw = cw.stft_tables(fs)["twiddle"]
    q1, j2 = np.meshgrid(np.arange(N // 32), np.arange(32), indexing="ij")
    want = np.exp(-2j * np.pi * (q1 * j2) / N) * (1j ** (q1 % 4))
    assert np.abs(tw - want).max() < 1e-6
    # and the decomposition equals a plain zero-padded FFT with the centring rotation
    rng = np.random.default_rng(3)
    u = (rng.standard_normal((4, L)) + 1j * rng.standard_normal((4, L))).astype(np.complex64)
    ref = np.fft.fft(u, n=N, axis=1) * (1j ** (np.arange(N) % 4))
    assert np.abs(fft_ax32(u, tw, N) - ref).max() < 1e-3


def test_stencil_properties(cw):
    for f in (-96000, -95818, -26000, -4400, 0, 37, 12345, 36000, 89636, 90000):
        c = cw.stft_channel(FS, f)
        assert c["q0"] % 2 == 0
        assert abs(abs(c["rot"]) - 1.0) < 1e-6
        w = c["wgt"]
        assert w.max() <= 1.0 + 1e-6 and w.min() >= 0.0 and w.argmax() in (3, 4)
    with pytest.raises(Exception):
        cw.stft_channel(FS, 90001)            # SSBD.hpp:102 "Signal outside of band (high)"
    with pytest.raises(Exception):
        cw.stft_channel(24000, 0)             # Fs/B must be an even integer >= 4 (SSBD.hpp:54)


@pytest.mark.parametrize("freq,usb", [(-96000, True), (-26000, True), (37, True), (89636, True), (26000, False)])
def test_stft_math_matches_the_oracle(cw, port, freq, usb):
    n = 40 * IQ_LEN
    sig = freq if usb else freq - 6000
    iq = synth.receiver_iq(n, FS, [sig], receiver=3, tones_per_channel=3)
    o = port.slot(FS, freq, iq, IQ_LEN, 0.9, af_size(15), is_usb=usb)
    wi = o["write_index"]
    got = stft_slot(cw, port, iq, freq, usb)[:wi]
    assert resid_db(got, o["raw"][:wi]) <= -110.0
    q = np.trunc(got * np.float32(o["factor"]) + np.float32(0.5)).astype(np.int32)
    assert np.abs(q - o["i16"][:wi].astype(np.int32)).max() <= 1


@pytest.mark.parametrize("fs,freq", [(96000, -48000), (96000, 12345), (96000, 42000), (48000, -24000), (48000, 1500),
                                     (48000, 18000)])
def test_stft_math_at_the_other_receiver_rates(cw, port, fs, freq):
    n = 40 * 1024
    iq = synth.receiver_iq(n, fs, [freq], receiver=4, tones_per_channel=3)
    o = port.slot(fs, freq, iq, 1024, 0.9, af_size(15))
    wi = o["write_index"]
    got = stft_slot(cw, port, iq, freq, True, fs)[:wi]
    assert resid_db(got, o["raw"][:wi]) <= -110.0
    q = np.trunc(got * np.float32(o["factor"]) + np.float32(0.5)).astype(np.int32)
    assert np.abs(q - o["i16"][:wi].astype(np.int32)).max() <= 1


@pytest.mark.parametrize("fs", [192000, 96000, 48000])
def test_work_items_cover_every_channel_with_its_own_stencil(cw, fs):
    """cwsl_stft_items (the grouping the channelizer kernel runs on): every channel sits in exactly one slot, its nine
    item weights are its own eight stencil weights at the same absolute bins and exact zeros elsewhere, the window
    starts on an even bin and a regular grid fills (nearly) all four slots -- for the stress sweep's regular grid, for
    random channel sets and for a grid far denser than the FFT's."""
    N, _, _ = geo(fs)
    half = fs // 2
    rng = np.random.default_rng(5)
    sets = {
        "stress": np.unique((-half + np.round(np.arange(1024) * (fs - 6000.0) / 1023)).astype(np.int32))[:1024],
        "random": rng.integers(-half, half - 6000, 300).astype(np.int32),
        "dense": (-20000 + 37 * np.arange(500)).astype(np.int32),
        "single": np.array([-26000 if fs == 192000 else -6000], np.int32),
    }
    for name, freqs in sets.items():
        usb = (np.arange(freqs.size) % 3 != 0).astype(np.int32) if name == "random" else np.ones(freqs.size, np.int32)
        legal = np.array([abs(int(f)) <= half and abs(int(f) + (6000 if u else -6000)) <= half for f, u in zip(freqs, usb)])
        freqs, usb = freqs[legal], usb[legal]
        it = cw.stft_items(fs, freqs, usb)
        members = it["channels"][it["channels"] >= 0]
        assert sorted(members.tolist()) == list(range(freqs.size)), name
        assert (it["first_bin"] % 2 == 0).all()
        for i in range(it["first_bin"].size):
            for j in range(4):
                c = it["channels"][i, j]
                if c < 0:
                    assert not it["weights"][i, j].any()
                    continue
                ch = cw.stft_channel(fs, int(freqs[c]), bool(usb[c]))
                want = np.zeros(12 + 16, np.float32)          # absolute bins first_bin-8 ... first_bin+19
                d = (ch["q0"] - it["first_bin"][i]) % N
                d = d - N if d > N // 2 else d
                want[8 + d: 8 + d + 8] = ch["wgt"]
                got = np.zeros_like(want)
                got[8 + it["shift"][j]: 8 + it["shift"][j] + 9] = it["weights"][i, j]
                assert np.array_equal(got, want), (name, i, j)
        if name == "stress" and fs == 192000:              # 0.97 bins between neighbours: every slot is used
            assert it["first_bin"].size == freqs.size // 4 == 256   # = one launch of the kernel
        if name == "single":
            assert it["first_bin"].size == 1
