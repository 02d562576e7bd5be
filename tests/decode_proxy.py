"""Modulation-level stand-in for the decoders the audio is handed to (TEST INFRASTRUCTURE, not a decoder).

north_star's last gate -- "the jt9/wsprd decode set must be identical" -- cannot be executed here: there is no
jt9/wsprd in the image and the reference does not contain one (it spawns the WSJT-X binaries,
source/DecoderPool.hpp:1007-1026). What CAN be executed is everything a decoder does before its channel code: find the
signal and turn the audio into per-symbol tone decisions and soft metrics. This module builds FT8-, FT4-, JT65- and WSPR-SHAPED
signals -- FT8: 8-FSK, 6.25 Hz tone spacing, 160 ms symbols, 79 symbols with the 7x7 Costas array 3,1,4,0,6,5,2 at symbols
0, 36 and 72; FT4: 4-FSK, 20.83 Hz, 48 ms symbols, four 4x4 Costas arrays; WSPR: 4-FSK, 1.46 Hz, 0.683 s symbols, 162
symbols with a sync bit in every symbol (the air-interface numbers printed in the WSJT-X user guide); the payload symbols
are random, NOT codewords -- and a non-coherent demodulator for them (Costas sync search over time/frequency, per-symbol tone energies from one
1920-point DFT per symbol, hard decisions, normalised soft metrics). The parity tests then require that the oracle's
int16 audio and the GPU's int16 audio give the SAME sync position, the SAME 79 hard decisions and soft metrics that
agree to 1e-3, at SNRs from far above to below the FT8 decoding threshold. Identical demodulator output means an
identical decoder input; it is evidence for, not a proof of, identical decode sets.
"""
from __future__ import annotations

import numpy as np

COSTAS = (3, 1, 4, 0, 6, 5, 2)
AUDIO_SR = 12000


class Waveform:
    """An M-FSK air interface as far as a demodulator sees it: tone count and spacing, symbol length at 12 kHz, frame
    length, and where which sync tones sit ({symbol index: tone})."""

    def __init__(self, name, ntones, nsps, nsym, sync):
        self.name, self.ntones, self.nsps, self.nsym, self.sync = name, ntones, nsps, nsym, dict(sync)
        self.tone_hz = AUDIO_SR / nsps          # orthogonal spacing: the tones sit on the bins of one nsps-point DFT

    def symbols(self, rng: np.random.Generator) -> np.ndarray:
        sym = rng.integers(0, self.ntones, self.nsym)
        for k, t in self.sync.items():
            sym[k] = t
        return sym.astype(np.int64)


def _costas_at(starts, pattern):
    return {s + k: c for s in starts for k, c in enumerate(pattern)}


# FT8: 8-FSK, 6.25 Hz, 160 ms symbols, 79 symbols, 7x7 Costas array at symbols 0, 36 and 72.
FT8 = Waveform("FT8-shaped", 8, 1920, 79, _costas_at((0, 36, 72), COSTAS))
# FT4: 4-FSK, 20.833 Hz, 48 ms symbols, 103 symbols between the ramp symbols, four different 4x4 Costas arrays.
FT4 = Waveform("FT4-shaped", 4, 576, 103, {**_costas_at((0,), (0, 1, 3, 2)), **_costas_at((33,), (1, 0, 2, 3)),
                                            **_costas_at((66,), (2, 3, 1, 0)), **_costas_at((99,), (3, 2, 0, 1))})
# WSPR: 4-FSK, 1.4648 Hz, 0.683 s symbols, 162 symbols; every symbol carries one sync bit in its LSB. The real sync
# vector is a fixed pseudo-random sequence; a stand-in of the same kind is drawn here (its values do not matter to a
# demodulator comparison), so only the tones {0,1} / {2,3} alternatives are constrained: modelled as 81 known symbols.
_wspr_rng = np.random.default_rng(162)
WSPR = Waveform("WSPR-shaped", 4, 8192, 162, {int(k): int(t) for k, t in zip(range(0, 162, 2), _wspr_rng.integers(0, 4, 81))})
# JT65: 65-FSK + a sync tone, 2.69 Hz tone spacing, 0.372 s symbols (4096 samples at 11025 Hz = 4458 at 12 kHz; 4456 here, a multiple of 8), 126
# symbols in a 60 s slot; the sync tone (tone 0) occupies 63 pseudo-randomly placed symbols, data tones start two
# spacings above it. As for WSPR the positions are a stand-in of the same kind as the real pseudo-random vector.
_jt65_rng = np.random.default_rng(65)
JT65 = Waveform("JT65-shaped", 67, 4456, 126, {int(k): 0 for k in np.sort(_jt65_rng.choice(126, 63, replace=False))})
_jt65_symbols = JT65.symbols


def _jt65_data_symbols(rng):  # data symbols never use tone 0 / 1 (sync tone and its guard spacing)
    sym = _jt65_symbols(rng)
    data = np.array([k not in JT65.sync for k in range(JT65.nsym)])
    sym[data] = rng.integers(2, JT65.ntones, int(data.sum()))
    return sym


JT65.symbols = _jt65_data_symbols
NSYM, NSPS, TONE_HZ = FT8.nsym, FT8.nsps, FT8.tone_hz
SYNC_AT = (0, 36, 72)


def make_symbols(rng: np.random.Generator, wf: Waveform = FT8) -> np.ndarray:
    return wf.symbols(rng)


def fsk_iq(n: int, fs: int, rf_hz: float, t0_s: float, symbols: np.ndarray, amplitude: float, wf: Waveform = FT8) -> np.ndarray:
    """(n,) complex128: continuous-phase M-FSK whose tone 0 sits rf_hz above the receiver's LO, starting t0_s into the
    slot (zero before and after)."""
    sps = int(round(fs * wf.nsps / AUDIO_SR))
    i0 = int(round(t0_s * fs))
    f = np.zeros(n, np.float64)
    on = np.zeros(n, bool)
    for k, s in enumerate(symbols):
        a, b = i0 + k * sps, min(n, i0 + (k + 1) * sps)
        if a >= n:
            break
        f[a:b] = rf_hz + wf.tone_hz * float(s)
        on[a:b] = True
    ph = 2.0 * np.pi * np.cumsum(f) / fs
    return np.where(on, amplitude * np.exp(1j * ph), 0.0)


def amplitude_for_snr(snr_db_2500: float, sigma: float, fs: int) -> float:
    """Amplitude of a complex exponential that is snr_db above complex white noise (sigma per component) measured in
    2500 Hz -- the bandwidth WSJT-X quotes its SNRs in."""
    noise = 2.0 * sigma * sigma * 2500.0 / fs
    return float(np.sqrt(noise * 10.0 ** (snr_db_2500 / 10.0)))


def tone_energies(audio: np.ndarray, f0_hz: float, start: int, wf: Waveform = FT8) -> np.ndarray:
    """(nsym, ntones) energies of the tones in every symbol window starting at sample `start` (windows that leave the
    buffer count as silence). f0_hz must be a multiple of the tone spacing: the tones then sit on DFT bins."""
    x = np.zeros(wf.nsym * wf.nsps, np.float64)
    a, b = max(0, start), min(audio.size, start + wf.nsym * wf.nsps)
    if b > a:
        x[a - start:b - start] = audio[a:b]
    spec = np.fft.rfft(x.reshape(wf.nsym, wf.nsps), axis=1)
    k0 = int(round(f0_hz / wf.tone_hz))
    return np.abs(spec[:, k0:k0 + wf.ntones]) ** 2


def sync_search(audio: np.ndarray, f0_hz: float, start: int, wf: Waveform = FT8, dt=None, df=range(-3, 4)):
    """Sync-tone correlation over a grid of time / frequency offsets around the nominal position: returns the best
    (dt_samples, df_bins) and the whole metric grid."""
    if dt is None:
        dt = range(-wf.nsps // 2, wf.nsps // 2 + 1, wf.nsps // 8)
    ks = np.array(sorted(wf.sync))
    ts = np.array([wf.sync[k] for k in ks])
    grid = np.zeros((len(dt), len(df)))
    for i, d in enumerate(dt):
        for j, q in enumerate(df):
            e = tone_energies(audio, f0_hz + q * wf.tone_hz, start + d, wf)
            grid[i, j] = float((e[ks, ts] / (e[ks].sum(axis=1) + 1e-30)).sum())
    i, j = np.unravel_index(int(grid.argmax()), grid.shape)
    return (list(dt)[i], list(df)[j]), grid


def demodulate(audio_i16: np.ndarray, f0_hz: float, t0_s: float, wf: Waveform = FT8):
    """What a decoder's front half extracts from one signal: sync position, hard tone decisions, soft metrics (each
    symbol's tone energies normalised to their sum)."""
    audio = audio_i16.astype(np.float64)
    start = int(round(t0_s * AUDIO_SR))
    pos, grid = sync_search(audio, f0_hz, start, wf)
    e = tone_energies(audio, f0_hz + pos[1] * wf.tone_hz, start + pos[0], wf)
    soft = e / (e.sum(axis=1, keepdims=True) + 1e-30)
    return dict(sync=pos, sync_grid=grid, hard=e.argmax(axis=1), soft=soft)
