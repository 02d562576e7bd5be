"""Numerical prototype (numpy, CPU) of the STFT channelizer form of the receive chain, checked against the
oracle: per hop of 16 IQ samples one 1024-point FFT of the (deconvolved-)windowed last 512 samples, then per
channel a w-tap Kaiser-Bessel interpolation between bins at the channel's exact NCO frequency, times the
reference's own drifting NCO phase.  Decides kernel width / accuracy before any CUDA is written."""
import os
import sys
import time

import numpy as np
import scipy.fft
import scipy.special

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cwsl_digi_b200 import synth          # noqa: E402
from oracle import oracle                  # noqa: E402

FS, IQ_LEN, N, L, HOP = 192000, 2048, 1024, 512, 16


def kb_kernel(w, sigma=2.0):
    beta = np.pi * np.sqrt((w / sigma) ** 2 * (sigma - 0.5) ** 2 - 0.8)

    def psi(t):                     # |t| <= w/2
        a = 1.0 - (2.0 * t / w) ** 2
        return np.where(a >= 0, scipy.special.i0(beta * np.sqrt(np.maximum(a, 0))) / scipy.special.i0(beta), 0.0)

    def psihat(s):                  # continuous FT of psi at frequency s (cycles per bin)
        z = np.sqrt((beta ** 2 - (np.pi * w * s) ** 2).astype(complex))
        return np.real(w * np.sinh(z) / z) / scipy.special.i0(beta)
    return psi, psihat


def channelize(iq, freqs, port, w, f32=True):
    n = iq.size // 2
    x = (iq[0::2] + 1j * iq[1::2]).astype(np.complex64 if f32 else np.complex128)
    nb = n // HOP
    xpad = np.concatenate([np.zeros(L - HOP, x.dtype), x])
    frames = np.lib.stride_tricks.sliding_window_view(xpad, L)[::HOP][:nb]
    psi, psihat = kb_kernel(w)
    t0 = port.tables(FS, int(freqs[0]))
    h = t0["filter"].astype(np.float64)
    j = np.arange(L)
    hd = h / psihat((j - L // 2) / N)
    rdt = np.float32 if f32 else np.float64
    hd = hd.astype(rdt)
    out = {}
    CH = 4096
    spec = np.empty((nb, N), x.dtype)
    iq_rot = (1j ** (np.arange(N) % 4)).astype(x.dtype)
    for b0 in range(0, nb, CH):
        u = frames[b0:b0 + CH] * hd
        spec[b0:b0 + CH] = scipy.fft.fft(u, n=N, axis=1) * iq_rot
    for f in freqs:
        tb = port.tables(FS, int(f))
        pinc = tb["phase_inc"].astype(np.float64)
        theta = np.arctan2(pinc[1], pinc[0])
        delta = -2.0 * np.pi * (f + 3000.0) / FS                       # SSBD.hpp:111 (USB)
        theta += 2.0 * np.pi * np.round((HOP * delta - theta) / (2.0 * np.pi))   # unwrap the per-block angle
        omega = theta / HOP
        nu = (-omega * N / (2 * np.pi)) % N
        q0 = int(np.ceil(nu - w / 2))
        qs = q0 + np.arange(w)
        wg = psi(nu - qs).astype(rdt)
        ph = port.phase_table(tb["phase_inc"], nb)
        ph = ph[:, 0].astype(np.float64) + 1j * ph[:, 1].astype(np.float64)
        rot = np.exp(-1j * omega * (L - HOP - L // 2))
        acc = (spec[:, qs % N] * wg).sum(axis=1)
        y = acc * (ph * rot).astype(x.dtype)
        b = np.arange(nb) & 3
        sign = 1.0
        a = np.where(b == 0, y.real, np.where(b == 1, -y.imag * sign, np.where(b == 2, -y.real, y.imag * sign)))
        out[int(f)] = a.astype(np.float32)
    return out


def resid_db(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return 10 * np.log10((d * d).sum() / max((b.astype(np.float64) ** 2).sum(), 1e-300))


def main():
    secs = float(os.environ.get("SECS", "1.0"))
    nblk = int(secs * FS / IQ_LEN)
    n = nblk * IQ_LEN
    freqs = [-96000, -95818, -26000, -4400, 0, 37, 12345, 36000, 89636, 90000]
    iq = synth.receiver_iq(n, FS, freqs, receiver=0)
    port = oracle.Port()
    afs = oracle.af_size(15.0)
    refs = {f: port.slot(FS, f, iq, IQ_LEN, 0.9, afs) for f in freqs}
    for w in (5, 6, 7, 8):
        for f32 in (False, True):
            t = time.time()
            got = channelize(iq, freqs, port, w, f32)
            line = []
            worst_lsb = 0
            for f in freqs:
                o = refs[f]
                wi = o["write_index"]
                r = resid_db(got[f][:wi], o["raw"][:wi])
                q = np.trunc(got[f][:wi] * np.float32(o["factor"]) + np.float32(0.5)).astype(np.int32)
                worst_lsb = max(worst_lsb, int(np.abs(q - o["i16"][:wi].astype(np.int32)).max()))
                line.append(f"{r:7.1f}")
            print(f"w={w} {'f32' if f32 else 'f64'}: resid dB per channel {' '.join(line)}  max LSB diff {worst_lsb}  ({time.time() - t:.1f}s)")


if __name__ == "__main__":
    main()
