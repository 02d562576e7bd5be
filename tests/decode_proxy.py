"""Modulation-level stand-in for the decoders the audio is handed to (TEST INFRASTRUCTURE, not a decoder).

north_star's last gate -- "the jt9/wsprd decode set must be identical" -- cannot be executed here: there is no
jt9/wsprd in the image and the reference does not contain one (it spawns the WSJT-X binaries,
source/DecoderPool.hpp:1007-1026). What CAN be executed is everything a decoder does before its channel code: find the
signal and turn the audio into per-symbol tone decisions and soft metrics. This module builds FT8-SHAPED signals --
8-FSK, 6.25 Hz tone spacing, 160 ms symbols, 79 symbols with the 7x7 Costas array 3,1,4,0,6,5,2 at symbols 0, 36 and 72
(the air-interface numbers printed in the WSJT-X user guide); the 58 payload symbols are random, NOT LDPC codewords -- and
a non-coherent demodulator for them (Costas sync search over time/frequency, per-symbol tone energies from one
1920-point DFT per symbol, hard decisions, normalised soft metrics). The parity tests then require that the oracle's
int16 audio and the GPU's int16 audio give the SAME sync position, the SAME 79 hard decisions and soft metrics that
agree to 1e-3, at SNRs from far above to below the FT8 decoding threshold. Identical demodulator output means an
identical decoder input; it is evidence for, not a proof of, identical decode sets.
"""
from __future__ import annotations

import numpy as np

COSTAS = (3, 1, 4, 0, 6, 5, 2)
NSYM, NSPS, TONE_HZ, AUDIO_SR = 79, 1920, 6.25, 12000     # 79 symbols of 1920 samples at 12 kHz: 12.64 s
SYNC_AT = (0, 36, 72)


def make_symbols(rng: np.random.Generator) -> np.ndarray:
    sym = rng.integers(0, 8, NSYM)
    for s in SYNC_AT:
        sym[s:s + 7] = COSTAS
    return sym.astype(np.int64)


def fsk_iq(n: int, fs: int, rf_hz: float, t0_s: float, symbols: np.ndarray, amplitude: float) -> np.ndarray:
    """(n,) complex128: continuous-phase 8-FSK whose tone 0 sits rf_hz above the receiver's LO, starting t0_s into the
    slot (zero before and after)."""
    sps = int(round(fs * NSPS / AUDIO_SR))
    i0 = int(round(t0_s * fs))
    f = np.zeros(n, np.float64)
    on = np.zeros(n, bool)
    for k, s in enumerate(symbols):
        a, b = i0 + k * sps, min(n, i0 + (k + 1) * sps)
        if a >= n:
            break
        f[a:b] = rf_hz + TONE_HZ * float(s)
        on[a:b] = True
    ph = 2.0 * np.pi * np.cumsum(f) / fs
    return np.where(on, amplitude * np.exp(1j * ph), 0.0)


def amplitude_for_snr(snr_db_2500: float, sigma: float, fs: int) -> float:
    """Amplitude of a complex exponential that is snr_db above complex white noise (sigma per component) measured in
    2500 Hz -- the bandwidth WSJT-X quotes its SNRs in."""
    noise = 2.0 * sigma * sigma * 2500.0 / fs
    return float(np.sqrt(noise * 10.0 ** (snr_db_2500 / 10.0)))


def tone_energies(audio: np.ndarray, f0_hz: float, start: int) -> np.ndarray:
    """(NSYM, 8) energies of the 8 tones in every symbol window starting at sample `start` (windows that leave the
    buffer count as silence). f0_hz must be a multiple of 6.25 Hz: the tones then sit on DFT bins."""
    x = np.zeros(NSYM * NSPS, np.float64)
    a, b = max(0, start), min(audio.size, start + NSYM * NSPS)
    if b > a:
        x[a - start:b - start] = audio[a:b]
    spec = np.fft.rfft(x.reshape(NSYM, NSPS), axis=1)
    k0 = int(round(f0_hz / TONE_HZ))
    return np.abs(spec[:, k0:k0 + 8]) ** 2


def sync_search(audio: np.ndarray, f0_hz: float, start: int, dt=range(-960, 961, 240), df=range(-3, 4)):
    """Costas correlation over a grid of time / frequency offsets around the nominal position: returns the best
    (dt_samples, df_bins) and the whole metric grid."""
    grid = np.zeros((len(dt), len(df)))
    for i, d in enumerate(dt):
        for j, q in enumerate(df):
            e = tone_energies(audio, f0_hz + q * TONE_HZ, start + d)
            m = 0.0
            for s in SYNC_AT:
                for k, c in enumerate(COSTAS):
                    m += e[s + k, c] / (e[s + k].sum() + 1e-30)
            grid[i, j] = m
    i, j = np.unravel_index(int(grid.argmax()), grid.shape)
    return (list(dt)[i], list(df)[j]), grid


def demodulate(audio_i16: np.ndarray, f0_hz: float, t0_s: float):
    """What a decoder's front half extracts from one signal: sync position, 79 hard tone decisions, soft metrics
    (each symbol's tone energies normalised to their sum)."""
    audio = audio_i16.astype(np.float64)
    start = int(round(t0_s * AUDIO_SR))
    pos, grid = sync_search(audio, f0_hz, start)
    e = tone_energies(audio, f0_hz + pos[1] * TONE_HZ, start + pos[0])
    soft = e / (e.sum(axis=1, keepdims=True) + 1e-30)
    return dict(sync=pos, sync_grid=grid, hard=e.argmax(axis=1), soft=soft)
