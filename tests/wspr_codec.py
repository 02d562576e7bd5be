"""WSPR channel code and a wsprd-style decoder (TEST INFRASTRUCTURE: the executable part of north_star's decode gate).

The reference hands every WSPR slot to the external ``wsprd`` (source/DecoderPool.hpp:1007-1026); neither wsprd nor
its sources are in the image, so the gate "decode set identical to the reference's" cannot run the real decoder. WSPR,
unlike FT8/FT4 (whose LDPC(174,91) generator cannot be derived offline), is fully specified by a handful of published
constants, and they can be VALIDATED here: the encoder below -- callsign/locator/power packing, the K = 32, r = 1/2
convolutional code with polynomials 0xF2D05351 / 0xE4613C47, the bit-reversal interleaver and the 162-bit sync vector
-- reproduces, symbol for symbol, the published channel symbols of the protocol's standard example "K1ABC FN42 37"
(``KAT_SYMBOLS``, the table printed in the WSPR protocol description; its LSBs are the sync vector, which pins that
table as well). Algorithm and answer were written down independently; 162 of 162 agree (tests/test_wspr_cpu.py).

The decoder follows wsprd's structure: 12 kHz int16 audio -> 375 Hz complex baseband around 1500 Hz (one long FFT, the
bins +-187.5 Hz inverse-transformed), half-symbol spectra, candidate search in the averaged spectrum, coarse sync
search over lag and frequency, fine search with matched per-symbol tone filters, soft symbols normalised to their
standard deviation, de-interleaving, sequential decoding of the K = 32 code (a stack decoder with the Fano metric and a
node budget), unpacking and a plausibility check of the message. It is an honest blind decoder: it is given the audio
and nothing else. Its sensitivity (about -28 dB in 2500 Hz; wsprd: -29 ... -31 dB) is what matters here, not its
pedigree: the gate compares DECODE SETS (message, frequency bin, time lag) obtained from the reference chain's audio
and from the GPU's audio, with signals on both sides of that threshold.
"""
from __future__ import annotations

import heapq

import numpy as np

AUDIO_SR = 12000
NSYM = 162
NSPS = 8192                      # samples per symbol at 12 kHz: 0.6827 s, tone spacing 12000 / 8192 = 1.4648 Hz
TONE_HZ = AUDIO_SR / NSPS
POLY1, POLY2 = 0xF2D05351, 0xE4613C47
NBITS = 81                       # 50 message bits + 31 zero tail bits
SYNC = np.array([
    1, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 0, 0, 0, 1, 0, 0, 1, 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 1,
    0, 0, 0, 0, 0, 0, 1, 0, 1, 1, 0, 0, 1, 1, 0, 1, 0, 0, 0, 1, 1, 0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 0, 1,
    0, 0, 1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 1, 1, 0, 1, 1, 0, 0, 1, 1,
    0, 1, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 1, 0, 1, 1, 0, 0, 0, 1, 1, 0,
    0, 0], np.int64)
# published channel symbols of "K1ABC FN42 37" (sync + 2 * data): the known-answer test of the whole encoder
KAT_MESSAGE = ("K1ABC", "FN42", 37)
KAT_SYMBOLS = np.array([int(c) for c in (
    "330020001020131222100323133220200032012322002232110233210221321222033030301210212032132003323032203020201023"
    "021112330231212221332000010320132222202332323320031222")], np.int64)
_INTERLEAVE = np.array([j for j in (int(f"{i:08b}"[::-1], 2) for i in range(256)) if j < NSYM], np.int64)
VALID_DBM = tuple(d for d in range(0, 61) if d % 10 in (0, 3, 7))


def _char_code(ch: str) -> int:
    if ch.isdigit():
        return ord(ch) - 48
    if ch == " ":
        return 36
    if "A" <= ch <= "Z":
        return ord(ch) - 65 + 10
    raise ValueError(f"bad callsign character {ch!r}")


def pack(call: str, grid: str, dbm: int) -> tuple[int, int]:
    """(N, M): the 28-bit callsign word and the 22-bit locator/power word of a type-1 message."""
    call = call.upper().strip()
    if len(call) >= 2 and call[1].isdigit() and not (len(call) >= 3 and call[2].isdigit()):
        call = " " + call                      # the third character is always the digit
    call = call.ljust(6)
    if len(call) != 6 or not call[2].isdigit() or any(c.isdigit() for c in call[3:]):
        raise ValueError(f"callsign {call!r} does not fit a type-1 message")
    c = [_char_code(x) for x in call]
    n = c[0]
    n = n * 36 + c[1]
    n = n * 10 + c[2]
    for k in (3, 4, 5):
        n = n * 27 + (c[k] - 10)               # letters 0..25, space 26
    grid = grid.upper()
    if len(grid) != 4 or not ("A" <= grid[0] <= "R" and "A" <= grid[1] <= "R" and grid[2:].isdigit()):
        raise ValueError(f"bad locator {grid!r}")
    m1 = (179 - 10 * (ord(grid[0]) - 65) - int(grid[2])) * 180 + 10 * (ord(grid[1]) - 65) + int(grid[3])
    if dbm not in VALID_DBM:
        raise ValueError(f"power {dbm} dBm is not a WSPR power level")
    return n, m1 * 128 + dbm + 64


def conv_encode(bits) -> np.ndarray:
    reg, out = 0, []
    for b in bits:
        reg = ((reg << 1) | int(b)) & 0xFFFFFFFF
        out.append((reg & POLY1).bit_count() & 1)
        out.append((reg & POLY2).bit_count() & 1)
    return np.array(out, np.int64)


def encode(call: str, grid: str, dbm: int) -> np.ndarray:
    """The 162 channel symbols (0..3) of a type-1 WSPR message."""
    n, m = pack(call, grid, dbm)
    bits = [(n >> (27 - i)) & 1 for i in range(28)] + [(m >> (21 - i)) & 1 for i in range(22)] + [0] * 31
    coded = conv_encode(bits)
    data = np.zeros(NSYM, np.int64)
    data[_INTERLEAVE] = coded
    return SYNC + 2 * data


def unpack(n: int, m: int):
    """Type-1 message from its two words, or None when they do not describe one."""
    ntype = (m & 127) - 64
    if ntype not in VALID_DBM or n >= 262177560:          # 37*36*10*27*27*27
        return None
    c = [0] * 6
    for k in (5, 4, 3):
        c[k] = n % 27 + 10
        n //= 27
    c[2] = n % 10
    n //= 10
    c[1] = n % 36
    n //= 36
    c[0] = n
    if c[0] > 36:
        return None
    call = "".join(" " if v == 36 else chr(48 + v) if v < 10 else chr(55 + v) for v in c).strip()
    m1 = m >> 7
    if m1 >= 32400:
        return None
    a, b = divmod(m1, 180)
    a = 179 - a
    grid = chr(65 + a // 10) + chr(65 + b // 10) + str(a % 10) + str(b % 10)
    if " " in call or not call or not ("A" <= grid[0] <= "R" and "A" <= grid[1] <= "R"):
        return None
    return call, grid, ntype


def fsk_audio_phase(symbols, fs: int, f0_hz: float, t0_s: float, n: int):
    """Instantaneous frequency track of a continuous-phase 4-FSK WSPR signal whose tone 0 sits at f0_hz: returns
    (phase[n] in radians, on[n])."""
    sps = int(round(fs * NSPS / AUDIO_SR))
    i0 = int(round(t0_s * fs))
    f = np.zeros(n, np.float64)
    on = np.zeros(n, bool)
    for k, s in enumerate(symbols):
        a, b = i0 + k * sps, min(n, i0 + (k + 1) * sps)
        if a >= n:
            break
        f[a:b] = f0_hz + TONE_HZ * float(s)
        on[a:b] = True
    return 2.0 * np.pi * np.cumsum(f) / fs, on


# ------------------------------------------------------------------------------------------------------------------
# decoder
# ------------------------------------------------------------------------------------------------------------------
BB_SR = 375                       # baseband rate: 12000 / 32
BB_NSPS = 256                     # samples per symbol at 375 Hz
_NFFT_IN = 46080 * 32             # 122.88 s at 12 kHz
_DF = BB_SR / 512.0               # half-symbol spectra: 0.732 Hz per bin = half the tone spacing


def to_baseband(audio_i16: np.ndarray) -> np.ndarray:
    """Complex baseband at 375 Hz centred on 1500 Hz audio (46080 samples = 122.88 s)."""
    x = np.zeros(_NFFT_IN, np.float64)
    n = min(_NFFT_IN, audio_i16.size)
    x[:n] = audio_i16[:n]
    spec = np.fft.rfft(x)
    k0 = int(round(1500.0 * _NFFT_IN / AUDIO_SR))
    half = 46080 // 2
    sel = np.concatenate([spec[k0:k0 + half], spec[k0 - half:k0]])    # fftshifted back: [0 .. +fmax, -fmax .. -0]
    return np.fft.ifft(sel) * (46080.0 / _NFFT_IN) * 2.0


def _half_symbol_spectra(bb: np.ndarray):
    """|FFT|^2 of 512-sample sine-windowed frames every 128 samples (half a symbol): ps[frame, bin], bins fftshifted
    so that bin 256 is 0 Hz (1500 Hz audio)."""
    nfr = (bb.size - 512) // 128 + 1
    idx = np.arange(512)[None, :] + 128 * np.arange(nfr)[:, None]
    w = np.sin(np.pi * np.arange(512) / 512.0)
    ps = np.abs(np.fft.fftshift(np.fft.fft(bb[idx] * w, axis=1), axes=1)) ** 2
    return ps


def _candidates(ps: np.ndarray, max_cand: int):
    """Peaks of the smoothed average spectrum inside +-110 Hz, as bins of the half-symbol spectra, strongest first."""
    avg = ps.mean(axis=0)
    sm = np.convolve(avg, np.ones(7), mode="same")                 # a signal occupies 4 tones = 7 bins
    lo, hi = 256 - int(110 / _DF), 256 + int(110 / _DF)
    band = sm[lo:hi + 1]
    noise = np.sort(band)[int(0.3 * band.size)]                     # the 30th percentile stands for the noise level
    rel = band / noise - 1.0
    cands = []
    for i in range(1, rel.size - 1):
        if rel[i] > rel[i - 1] and rel[i] >= rel[i + 1] and rel[i] > 0.02:
            cands.append((float(rel[i]), lo + i))
    cands.sort(reverse=True)
    return [b for _, b in cands[:max_cand]]


def _coarse_sync(ps: np.ndarray, bin0: int):
    """Best (score, lag_frames, bin) around a candidate: sync-vector correlation on amplitudes, frames are half
    symbols, the four tones sit at bin-3, bin-1, bin+1, bin+3 of the 0.73 Hz grid."""
    amp = np.sqrt(ps)
    sgn = 2.0 * SYNC - 1.0
    best = (-1e30, 0, bin0)
    nfr = amp.shape[0]
    for b in range(bin0 - 2, bin0 + 3):
        if b - 3 < 0 or b + 3 >= amp.shape[1]:
            continue
        for lag in range(-8, 23):                                    # -2.7 ... +7.5 s around the nominal 1 s start
            fr = lag + 2 * np.arange(NSYM) + 3                       # nominal start 1 s = frame 2.93
            ok = (fr >= 0) & (fr < nfr)
            if ok.sum() < 120:
                continue
            f = fr[ok]
            p0, p1, p2, p3 = amp[f, b - 3], amp[f, b - 1], amp[f, b + 1], amp[f, b + 3]
            s = float((((p1 + p3) - (p0 + p2)) * sgn[ok]).sum() / (p0 + p1 + p2 + p3).sum())
            if s > best[0]:
                best = (s, lag, b)
    return best


def _symbol_amplitudes(bb: np.ndarray, f_hz: float, start: int) -> np.ndarray:
    """(162, 4) matched-filter amplitudes of the four tones (centre frequency f_hz relative to 1500 Hz) for symbols
    starting at baseband sample `start`."""
    n = NSYM * BB_NSPS
    seg = np.zeros(n, np.complex128)
    a, b = max(0, start), min(bb.size, start + n)
    if b > a:
        seg[a - start:b - start] = bb[a:b]
    t = np.arange(BB_NSPS) / BB_SR
    out = np.empty((NSYM, 4))
    blocks = seg.reshape(NSYM, BB_NSPS)
    for k in range(4):
        fk = f_hz + (k - 1.5) * TONE_HZ
        out[:, k] = np.abs(blocks @ np.exp(-2j * np.pi * fk * t))
    return out


def _sync_metric(p: np.ndarray) -> float:
    return float(((p[:, 1] + p[:, 3] - p[:, 0] - p[:, 2]) * (2.0 * SYNC - 1.0)).sum() / p.sum())


def _soft_symbols(p: np.ndarray) -> np.ndarray:
    """One soft value per symbol for its DATA bit, given the known sync bit: amplitude of the tone with data = 1 minus
    that with data = 0, normalised to unit standard deviation (wsprd's symfac step)."""
    s = np.where(SYNC == 1, p[:, 3] - p[:, 1], p[:, 2] - p[:, 0])
    s = s - s.mean()
    return s / (s.std() + 1e-30)


def _stack_decode(soft: np.ndarray, max_nodes: int):
    """Sequential (stack) decoding of the K = 32, r = 1/2 code with the Fano metric. soft: 162 values, positive = 1,
    unit variance, de-interleaved. Returns the 81 decoded bits or None when the node budget is spent."""
    # Fano metric per coded bit: log2(2 P(bit | x)) - R with P from a logistic model of the soft value
    a = 1.6
    l1 = np.log2(2.0 / (1.0 + np.exp(-a * soft))) - 0.5
    l0 = np.log2(2.0 / (1.0 + np.exp(+a * soft))) - 0.5
    met = (l0.tolist(), l1.tolist())
    heap = [(0.0, 0, 0, 0)]         # (-metric, -depth, encoder register, decoded bits): best metric first, deeper on ties
    nodes = 0
    while heap and nodes < max_nodes:
        negm, depth, reg, bits = heapq.heappop(heap)
        depth = -depth
        if depth == NBITS:
            return [(bits >> (NBITS - 1 - i)) & 1 for i in range(NBITS)]
        nodes += 1
        for b in ((0,) if depth >= 50 else (0, 1)):                  # the 31 tail bits are zeros
            r = ((reg << 1) | b) & 0xFFFFFFFF
            c1 = (r & POLY1).bit_count() & 1
            c2 = (r & POLY2).bit_count() & 1
            m = -negm + met[c1][2 * depth] + met[c2][2 * depth + 1]
            heapq.heappush(heap, (-m, -(depth + 1), r, (bits << 1) | b))
    return None


def decode(audio_i16: np.ndarray, max_cand: int = 24, max_nodes: int = 20000):
    """Blind decode of one WSPR slot: list of dicts (message, call, grid, dbm, freq_hz, dt_s, snr_db, sync), one per
    distinct message, strongest candidate first."""
    bb = to_baseband(np.asarray(audio_i16))
    ps = _half_symbol_spectra(bb)
    found, seen = [], set()
    noise_bin = np.sort(ps.mean(axis=0)[256 - 150:256 + 151])[90]    # 30th percentile of the average spectrum
    for bin0 in _candidates(ps, max_cand):
        score, lag, b = _coarse_sync(ps, bin0)
        if score < 0.10:
            continue
        f0 = (b - 256) * _DF
        start0 = (lag + 4) * 128                                     # frame fr is centred on sample 128 fr + 256
        best = (-1e30, f0, start0, None)
        for df in (-0.37, -0.18, 0.0, 0.18, 0.37):                   # fine search: +-half a coarse bin, +-3/4 frame
            for ds in (-96, -64, -32, 0, 32, 64, 96):
                p = _symbol_amplitudes(bb, f0 + df, start0 + ds)
                s = _sync_metric(p)
                if s > best[0]:
                    best = (s, f0 + df, start0 + ds, p)
        s, f, start, p = best
        soft = _soft_symbols(p)[_INTERLEAVE]                          # de-interleave: coded bit j came from symbol
        bits = _stack_decode(soft, max_nodes)
        if bits is None:
            continue
        n = int("".join(map(str, bits[:28])), 2)
        m = int("".join(map(str, bits[28:50])), 2)
        msg = unpack(n, m)
        if msg is None:
            continue
        # a decode must re-encode to symbols that the soft values mostly agree with (rejects budget-exhausted junk)
        again = encode(*msg)
        agree = float(((again >> 1) == (_soft_symbols(p) > 0)).mean())
        if agree < 0.62:
            continue
        if msg in seen:
            continue
        seen.add(msg)
        # SNR in 2500 Hz: tone power from the matched filters at the re-encoded symbols (noise floor removed) over
        # the noise the half-symbol spectra see per bin (sum of the sine window's squares = 256)
        sig = float((p[np.arange(NSYM), again] ** 2).mean()) / BB_NSPS ** 2 - noise_bin / 256.0 / BB_NSPS
        snr = 10.0 * np.log10(max(sig, 1e-30) / (noise_bin / 256.0 * 2500.0 / BB_SR))
        found.append(dict(message=f"{msg[0]} {msg[1]} {msg[2]}", call=msg[0], grid=msg[1], dbm=msg[2],
                          freq_hz=round(1500.0 + f, 2), dt_s=round(start / BB_SR - 1.0, 3), snr_db=round(float(snr), 1),
                          sync=round(s, 4), agree=round(agree, 3)))
    return found


def decode_set(audio_i16: np.ndarray, **kw):
    """What the gate compares: {(message, frequency to 0.01 Hz, lag to 1 ms)}."""
    return {(d["message"], d["freq_hz"], d["dt_s"]) for d in decode(audio_i16, **kw)}
