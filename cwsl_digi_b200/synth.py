"""Deterministic synthetic IQ for tests and bench.

Counter-based (SplitMix64 keyed by seed/receiver/sample index) so any sub-range can be produced
independently; Gaussian via Box-Muller in float64, rounded once to float32. Layout matches the
CWSL shared-memory blocks the reference consumes: interleaved float32 (I, Q),
``source/Receiver.hpp:140``.
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 20261017
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def gaussian_iq(n: int, receiver: int = 0, sigma: float = 300.0, seed: int = BASE_SEED,
                start: int = 0) -> np.ndarray:
    """(n, 2) float64 complex white Gaussian noise, sigma per component."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed + receiver) * np.uint64(0x100000001B3) + np.uint64(0x51ED27))
        a = _splitmix64(idx * np.uint64(2) + key)
        b = _splitmix64(idx * np.uint64(2) + np.uint64(1) + key)
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) / 9007199254740993.0   # (0,1)
    u2 = (b >> np.uint64(11)).astype(np.float64) / 9007199254740992.0           # [0,1)
    r = sigma * np.sqrt(-2.0 * np.log(u1))
    th = 2.0 * np.pi * u2
    return np.stack([r * np.cos(th), r * np.sin(th)], axis=1)


def tones(n: int, fs: int, freqs_hz, amplitude: float = 8000.0, start: int = 0) -> np.ndarray:
    """(n, 2) float64 sum of complex tones at baseband offsets ``freqs_hz`` (relative to the LO).
    Phase is computed from an exact integer (f*i mod fs) so long slots do not lose precision."""
    out = np.zeros((n, 2), np.float64)
    i = np.arange(start, start + n, dtype=np.int64)
    for j, f in enumerate(freqs_hz):
        fi = int(round(f))
        frac = float(f) - fi
        ph = 2.0 * np.pi * (((fi * i) % fs).astype(np.float64) / fs + frac * (i.astype(np.float64) / fs))
        ph += 0.61803398875 * j
        out[:, 0] += amplitude * np.cos(ph)
        out[:, 1] += amplitude * np.sin(ph)
    return out


def receiver_iq(n: int, fs: int, demod_freqs, receiver: int = 0, tones_per_channel: int = 8,
                sigma: float = 300.0, amplitude: float = 8000.0, seed: int = BASE_SEED) -> np.ndarray:
    """Interleaved float32 IQ (2n,) for one receiver: noise + ``tones_per_channel`` tones inside each
    decoder passband at audio offsets drawn uniformly in [200, 2900] Hz (RF = dial + offset)."""
    x = gaussian_iq(n, receiver, sigma, seed)
    rng_idx = np.arange(len(demod_freqs) * tones_per_channel, dtype=np.uint64)
    with np.errstate(over="ignore"):
        u = _splitmix64(rng_idx + np.uint64(seed + 7919 * (receiver + 1)))
    off = 200.0 + (u >> np.uint64(11)).astype(np.float64) / 9007199254740992.0 * 2700.0
    off = np.round(off)
    fl = []
    for c, f in enumerate(demod_freqs):
        for t in range(tones_per_channel):
            fl.append(float(f) + off[c * tones_per_channel + t])
    if fl:
        x = x + tones(n, fs, fl, amplitude)
    return np.ascontiguousarray(x, dtype=np.float32).reshape(-1)


def stress_demod_freqs(n_channels: int = 1024) -> np.ndarray:
    """BASELINE.json configs[4]: demodFreq_c = -96000 + round(c*186000/(n-1)) (legal USB range)."""
    if n_channels == 1:
        return np.array([-26000], np.int32)
    c = np.arange(n_channels, dtype=np.float64)
    return (-96000 + np.round(c * 186000.0 / (n_channels - 1))).astype(np.int32)
