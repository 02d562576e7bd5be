"""CPU check of the STFT channelizer's host-side tables (cwsl_stft_tables / cwsl_stft_channel, no device needed):
the kernel's arithmetic restated with numpy -- windowed 1024-point FFT per hop (32x32 decomposition with the
library's own twiddle table), 8-bin interpolation, phase table from the oracle -- must reproduce the oracle's audio
to the FAST-mode bars. Pins the math of cwsl_chan.cu on CPU; the kernel itself is checked by the -m gpu tests."""
import numpy as np
import pytest

from cwsl_digi_b200 import synth
from oracle.oracle import af_size

FS, IQ_LEN, N, L, HOP = 192000, 2048, 1024, 512, 16


def fft1024_32x32(u, tw):
    """X[q1 + 32 q2] of the zero-padded window, the way the kernel computes it (two 32-point passes + twiddles);
    includes the i^q centring rotation that the twiddle table carries."""
    nb = u.shape[0]
    up = np.zeros((nb, N), np.complex64)
    up[:, :L] = u
    a = up.reshape(nb, 32, 32)                       # [j1, j2]
    p1 = np.fft.fft(a, axis=1).astype(np.complex64)  # over j1 -> [q1, j2]
    p1 = p1 * tw[None, :, :]                         # W1024^(j2 q1) * i^q1
    p2 = np.fft.fft(p1, axis=2).astype(np.complex64)  # over j2 -> [q1, q2]
    return np.transpose(p2, (0, 2, 1)).reshape(nb, N)  # bin q1 + 32 q2


def stft_slot(cw, port, iq, freq, usb=True):
    x = (iq[0::2] + 1j * iq[1::2]).astype(np.complex64)
    nb = x.size // HOP
    t = cw.stft_tables(FS)
    xpad = np.concatenate([np.zeros(L - HOP, np.complex64), x])
    frames = np.lib.stride_tricks.sliding_window_view(xpad, L)[::HOP][:nb]
    spec = fft1024_32x32(frames * t["window"][None, :], t["twiddle"])
    c = cw.stft_channel(FS, freq, usb)
    acc = (spec[:, (c["q0"] + np.arange(8)) % N] * c["wgt"][None, :]).sum(axis=1)
    tb = port.tables(FS, freq, is_usb=usb)
    ph = port.phase_table(tb["phase_inc"], nb)
    y = acc * ((ph[:, 0] + 1j * ph[:, 1]) * c["rot"])
    sign = 1.0 if usb else -1.0
    b = np.arange(nb) & 3
    return np.where(b == 0, y.real, np.where(b == 1, -y.imag * sign, np.where(b == 2, -y.real, y.imag * sign))).astype(np.float32)


def resid_db(got, want):
    want = want.astype(np.float64)
    err = got.astype(np.float64) - want
    return 20 * np.log10(max(np.sqrt(np.mean(err ** 2)), 1e-300) / np.sqrt(np.mean(want ** 2)))


def test_twiddles_are_the_32x32_factors(cw):
    tw = cw.stft_tables(FS)["twiddle"]
    q1, j2 = np.meshgrid(np.arange(32), np.arange(32), indexing="ij")
    want = np.exp(-2j * np.pi * (q1 * j2) / N) * (1j ** (q1 % 4))
    assert np.abs(tw - want).max() < 1e-6
    # and the decomposition equals a plain zero-padded FFT with the centring rotation
    rng = np.random.default_rng(3)
    u = (rng.standard_normal((4, L)) + 1j * rng.standard_normal((4, L))).astype(np.complex64)
    ref = np.fft.fft(u, n=N, axis=1) * (1j ** (np.arange(N) % 4))
    assert np.abs(fft1024_32x32(u, tw) - ref).max() < 1e-3


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
        cw.stft_channel(96000, 0)             # built for 192 kHz receivers


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
