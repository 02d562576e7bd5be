"""CPU tests: the oracle restatement (oracle/cwsl_oracle.c) against the reference's own headers
(oracle/_ref), against the committed golden vectors, and against DSP facts that do not depend
on either."""
import glob
import os

import numpy as np
import pytest

from cwsl_digi_b200 import synth
from oracle.oracle import af_size

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ---- known answers on derived constants (measured on the reference build) ----
def test_getters_and_geometry(ref, port):
    g = ref.getters(192000)
    assert g == dict(InRate=192000, OutRate=12000, InSize=64, OutSize=4, Bandwidth=6000, Delay=8)
    t = port.tables(192000, -26000)
    assert (t["filt_order"], t["block_size"], t["num_ws"]) == (512, 16, 32)
    t = port.tables(96000, 0)
    assert (t["filt_order"], t["block_size"], t["num_ws"]) == (256, 8, 32)
    t = port.tables(48000, 0)
    assert (t["filt_order"], t["block_size"], t["num_ws"]) == (128, 4, 32)


def test_tap_anchors(port):
    t = port.tables(192000, -26000)
    f = t["filter"].astype(np.float64) * t["raw_tap_sum"]   # undo the DC normalisation
    assert abs(t["raw_tap_sum"] - 31.941248) < 1e-5
    assert f[0] == 0.0
    assert abs(f[256] - 1.0) < 1e-6
    assert abs(f[1] - (-0.000313357)) < 1e-8
    assert abs(f[255] - 0.99836) < 1e-5 and abs(f[257] - 0.99836) < 1e-5
    assert np.array_equal(t["filter"][1:], t["filter"][1:][::-1])        # f[n] == f[order-n]
    assert abs(float(np.sum(t["filter"].astype(np.float64))) - 1.0) < 1e-6


@pytest.mark.parametrize("fs", [192000, 96000, 48000])
@pytest.mark.parametrize("freq", [-24000, 0, 1, -4400, 12345, -7, 18000])
def test_tables_bit_identical_to_reference(ref, port, fs, freq):
    a, b = ref.tables(fs, freq), port.tables(fs, freq)
    for k in ("filter", "tone", "phase_inc"):
        assert np.array_equal(_bits(a[k]), _bits(b[k])), k
    assert a["raw_tap_sum"] == b["raw_tap_sum"]


def test_tuning_range_matches_reference(ref, port):
    # legal USB range at 192 kHz is [-96000, +90000] (SSBD.hpp:100-103)
    for f, ok in [(-96000, True), (-96001, False), (90000, True), (90001, False), (96000, False)]:
        for o in (ref, port):
            if ok:
                o.tables(192000, f)
            else:
                with pytest.raises(ValueError):
                    o.tables(192000, f)


def test_phase_recurrence_matches_reference(ref, port):
    t = port.tables(192000, -26000)
    n = 20000
    tab = port.phase_table(t["phase_inc"], n + 1)
    assert np.array_equal(_bits(tab[n]), _bits(ref.phase_after(192000, -26000, n)))
    assert tab[0, 0] == 1.0 and tab[0, 1] == 0.0


def test_phase_drift_is_the_references(port):
    # |phase| after 180000 float steps for F=-26000 drifts to 1.0004: the float recurrence is not a
    # pure rotation, which is why the product replays it instead of computing exact phases
    t = port.tables(192000, -26000)
    tab = port.phase_table(t["phase_inc"], 180001)
    assert abs(np.hypot(*tab[180000].astype(np.float64)) - 1.000401) < 2e-5


# ---- restatement == reference, bit for bit -----------------------------------------------------
@pytest.mark.parametrize("fs,iq_len,freq,scale,period", [
    (192000, 2048, -26000, 0.90, 15.0),
    (192000, 512, 90000, 0.20, 120.0),
    (192000, 4096, -96000, 0.90, 7.5),
    (96000, 1024, 30000, 0.90, 15.0),
    (48000, 512, -20000, 0.90, 15.0),
])
def test_port_equals_reference(ref, port, fs, iq_len, freq, scale, period):
    n = (fs // 2) // iq_len * iq_len
    iq = synth.receiver_iq(n, fs, [freq], receiver=3, tones_per_channel=3)
    afs = af_size(period)
    a = ref.slot(fs, freq, iq, iq_len, scale, afs)
    b = port.slot(fs, freq, iq, iq_len, scale, afs)
    assert a["write_index"] == b["write_index"] == n * 12000 // fs
    assert np.array_equal(_bits(a["raw"]), _bits(b["raw"]))
    assert np.array_equal(a["i16"], b["i16"])
    assert a["max"] == b["max"] and a["factor"] == b["factor"]


def test_af_buffer_full_guard(ref, port):
    # Instance.cpp:268-271: blocks are dropped once write_index + iq_len > size - 1
    fs, iq_len = 48000, 4096
    afs = 6000  # tiny buffer: accepts (6000-1-4096)/1024+1 = 2 blocks
    iq = synth.receiver_iq(5 * iq_len, fs, [0], tones_per_channel=1)
    a = ref.slot(fs, 0, iq, iq_len, 0.9, afs)
    b = port.slot(fs, 0, iq, iq_len, 0.9, afs)
    assert a["write_index"] == b["write_index"] == 2 * iq_len // 4
    assert np.array_equal(a["i16"], b["i16"])
    assert port.accepted_blocks(5 * iq_len, iq_len, 4, afs) == 2


def test_quantise_is_add_half_then_truncate(port):
    # Instance.cpp:238-241: (int16)(x + 0.5f): -3.7 -> -3, 3.7 -> 4 (asymmetric)
    import ctypes as C
    buf = np.array([-3.7, 3.7, -0.4, 0.49, -1.5, 1.5, 0.0], np.float32)
    out = np.zeros(buf.size, np.int16)
    port.lib.oracle_quantise.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    port.lib.oracle_quantise(buf.ctypes.data, buf.size, out.ctypes.data)
    assert out.tolist() == [-3, 4, 0, 0, -1, 2, 0]


# ---- golden vectors (generated from oracle/_ref by tests/golden/make_golden.py) -----------------
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_port_against_golden(port, path):
    g = np.load(path)
    fs, iq_len, afs = int(g["fs"]), int(g["iq_len"]), int(g["af_size"])
    for c, f in enumerate(g["freqs"]):
        o = port.slot(fs, int(f), g["iq"], iq_len, float(g["scales"][c]), afs)
        wi = int(g["write_index"][c])
        assert o["write_index"] == wi
        assert np.array_equal(_bits(o["raw"][:wi]), _bits(g["raw"][c]))
        assert np.array_equal(o["i16"][:wi], g["i16"][c]) and not o["i16"][wi:].any()
        assert o["max"] == g["maxval"][c] and o["factor"] == g["factor"][c]


def test_golden_present():
    assert len(GOLDEN) >= 4


# ---- functional DSP checks, independent of the oracle -----------------
def _tone_gain_db(port, offset_hz, fs=192000, dial=-26000):
    n = fs // 2
    i = np.arange(n)
    ph = 2 * np.pi * ((dial + offset_hz) * i % fs) / fs
    iq = np.stack([10000 * np.cos(ph), 10000 * np.sin(ph)], 1).astype(np.float32).reshape(-1)
    o = port.slot(fs, dial, iq, 2048, 0.9, af_size(15))
    a = o["raw"][200:o["write_index"]].astype(np.float64)
    return 20 * np.log10(np.sqrt(np.mean(a * a)) / (10000 / np.sqrt(2)))


@pytest.mark.parametrize("offset,lo,hi", [
    (1500, -0.05, 0.05), (2900, -0.05, 0.05),        # passband: unity gain
    (200, -2.6, -2.0), (5800, -2.6, -2.0),           # band edges
    (-500, -34.0, -31.5), (6500, -34.0, -31.5),      # transition band
    (-1500, -63.0, -59.0),                           # opposite sideband
    (-3500, -74.0, -69.0), (9000, -67.0, -62.0),     # stop band
])
def test_tone_response(port, offset, lo, hi):
    assert lo <= _tone_gain_db(port, offset) <= hi


def test_audio_tone_lands_at_offset(port):
    # a carrier at dial+1000 Hz must come out as a 1000 Hz audio tone (USB)
    fs, dial = 192000, 5000
    n = fs
    i = np.arange(n)
    ph = 2 * np.pi * ((dial + 1000) * i % fs) / fs
    iq = np.stack([8000 * np.cos(ph), 8000 * np.sin(ph)], 1).astype(np.float32).reshape(-1)
    o = port.slot(fs, dial, iq, 2048, 0.9, af_size(15))
    a = o["raw"][1000:1000 + 8192].astype(np.float64)
    spec = np.abs(np.fft.rfft(a * np.hanning(a.size)))
    assert abs(np.argmax(spec) * 12000 / a.size - 1000) < 3
