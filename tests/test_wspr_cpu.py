"""CPU tests of the WSPR codeword generator (host C++) and of the test-side WSPR decoder
that makes north_star's decode gate executable for WSPR (tests/wspr_codec.py).

* known-answer test: both encoders (Python here, C++ in cwsl_digi_b200/host/WsprSynth.hpp) reproduce the 162 published
  channel symbols of "K1ABC FN42 37"; they agree with each other on other messages;
* round trip: valid transmissions + noise -> the reference chain (oracle) -> blind decode gives the messages back with
  the right frequency, lag and SNR -- on IQ built here and on IQ produced by the C++ SyntheticIqSource."""
import os
import subprocess

import numpy as np
import pytest

import wspr_codec as wc
from cwsl_digi_b200 import synth
from oracle.oracle import af_size

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "cwsl_digi_b200", "host")
FS, IQ_LEN, SIGMA = 192000, 2048, 300.0
MESSAGES = [("K1ABC", "FN42", 37), ("W1AW", "FN31", 30), ("DL1ABC", "JO62", 23), ("G4JNT", "IO90", 10),
            ("VK3XYZ", "QF22", 0), ("JA1AA", "PM95", 60), ("N0ABC", "EM10", 33), ("F5ABC", "JN18", 27),
            ("9A1A", "JN75", 43), ("ZL1", "RF72", 7)]


def host_tests_exe():
    subprocess.run(["make", "-C", HOST, "host_tests"], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(HOST, "host_tests")


def test_known_answer():
    assert wc.KAT_SYMBOLS.size == 162 and np.array_equal(wc.KAT_SYMBOLS & 1, wc.SYNC)
    assert np.array_equal(wc.encode(*wc.KAT_MESSAGE), wc.KAT_SYMBOLS)


@pytest.mark.parametrize("msg", MESSAGES)
def test_pack_unpack_and_cpp_encoder(cw, msg):
    n, m = wc.pack(*msg)
    assert n < 1 << 28 and m < 1 << 22 and wc.unpack(n, m) == msg
    r = subprocess.run([host_tests_exe(), "wspr", msg[0], msg[1], str(msg[2])], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "".join(map(str, wc.encode(*msg)))


def test_random_messages_round_trip_and_match_cpp(cw):
    """Seeded random type-1 messages: pack/unpack round trip, every symbol's LSB is the sync vector, the code is linear
    in the message bits (conv_encode(a ^ b) = conv_encode(a) ^ conv_encode(b)), and the C++ encoder emits the same
    162 symbols."""
    rng = np.random.default_rng(20261017)
    letters = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"
    exe = host_tests_exe()
    prev = None
    for _ in range(24):
        c0 = rng.choice(list(letters + "0123456789 "))
        c1 = rng.choice(list(letters))
        suffix = "".join(rng.choice(list(letters), int(rng.integers(0, 4))))
        call = (c0 + c1 + str(int(rng.integers(0, 10))) + suffix).strip()
        grid = letters[rng.integers(0, 18)] + letters[rng.integers(0, 18)] + str(int(rng.integers(0, 10))) + str(int(rng.integers(0, 10)))
        dbm = int(rng.choice(wc.VALID_DBM))
        if call[1].isdigit() and not (len(call) > 2 and call[2].isdigit()):
            continue                                   # (a leading blank was stripped: the digit moved to second place)
        n, m = wc.pack(call, grid, dbm)
        assert wc.unpack(n, m) == (call, grid, dbm)
        sym = wc.encode(call, grid, dbm)
        assert np.array_equal(sym & 1, wc.SYNC) and sym.max() <= 3
        r = subprocess.run([exe, "wspr", call, grid, str(dbm)], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.strip() == "".join(map(str, sym)), (call, grid, dbm)
        bits = [(n >> (27 - i)) & 1 for i in range(28)] + [(m >> (21 - i)) & 1 for i in range(22)] + [0] * 31
        if prev is not None:
            x = [a ^ b for a, b in zip(bits, prev)]
            assert np.array_equal(wc.conv_encode(x), wc.conv_encode(bits) ^ wc.conv_encode(prev))
        prev = bits


def test_bad_messages_rejected(cw):
    for bad in [("N0CALL", "FN20", 30), ("K1ABC", "FN4", 37), ("K1ABC", "FN42", 35), ("K1AB1", "FN42", 37), ("K1ABC", "SS00", 37)]:
        with pytest.raises(ValueError):
            wc.encode(*bad)
        assert subprocess.run([host_tests_exe(), "wspr", bad[0], bad[1], str(bad[2])], capture_output=True).returncode == 3


def test_stack_decoder_corrects_errors():
    """The channel code alone: 162 hard symbols with 20 of the data bits flipped still decode to the message, and
    the 31 tail bits come back as zeros."""
    rng = np.random.default_rng(7)
    sym = wc.encode("DL1ABC", "JO62", 23)
    soft = np.where(sym >> 1 == 1, 1.0, -1.0)
    flip = rng.choice(162, 20, replace=False)
    soft[flip] *= -1
    bits = wc._stack_decode(soft[wc._INTERLEAVE] * 1.0, 50000)
    assert bits is not None
    n, m = int("".join(map(str, bits[:28])), 2), int("".join(map(str, bits[28:50])), 2)
    assert wc.unpack(n, m) == ("DL1ABC", "JO62", 23) and not any(bits[50:])


def wspr_slot_iq(plan, demod_hz, receiver=5, seconds=120.0):
    """Complex noise + the transmissions of `plan` [(call, grid, dbm, snr_db, audio_hz, t0_s)] around demod_hz."""
    n = int(seconds * FS) // IQ_LEN * IQ_LEN
    x = synth.gaussian_iq(n, receiver=receiver, sigma=SIGMA)
    z = x[:, 0] + 1j * x[:, 1]
    del x
    for call, grid, dbm, snr, fa, t0 in plan:
        ph, on = wc.fsk_audio_phase(wc.encode(call, grid, dbm), FS, demod_hz + fa - 1.5 * wc.TONE_HZ, t0, n)
        a = np.sqrt(2.0 * SIGMA ** 2 * 2500.0 / FS * 10.0 ** (snr / 10.0))
        z[on] += a * np.exp(1j * ph[on])
    return np.ascontiguousarray(np.stack([z.real, z.imag], axis=1), np.float32).reshape(-1)


def test_round_trip_through_the_reference_chain(port):
    plan = [("K1ABC", "FN42", 37, -12.0, 1500.0, 1.0), ("W1AW", "FN31", 30, -20.0, 1440.0, 1.3),
            ("G4JNT", "IO90", 23, -24.0, 1570.0, 0.8)]
    iq = wspr_slot_iq(plan, -4400)
    audio = port.slot(FS, -4400, iq, IQ_LEN, 0.2, af_size(120.0))["i16"]
    got = {d["message"]: d for d in wc.decode(audio)}
    assert set(got) == {f"{c} {g} {p}" for c, g, p, *_ in plan}
    for c, g, p, snr, fa, t0 in plan:
        d = got[f"{c} {g} {p}"]
        assert abs(d["freq_hz"] - fa) < 0.4 and abs(d["dt_s"] - (t0 - 1.0)) < 0.1 and abs(d["snr_db"] - snr) < 1.5, d


def test_cpp_source_is_decodable(cw, port, tmp_path):
    """SyntheticIqSource + makeWsprBurst (the generator station_demo uses): 120 s of IQ written by the C++ side, through
    the reference chain, decodes to the transmitted message."""
    out = tmp_path / "wspr.f32"
    subprocess.run([host_tests_exe(), "wspriq", str(out), "120", "-18"], check=True)
    iq = np.fromfile(out, np.float32)
    assert iq.size == int(120 * FS) // IQ_LEN * IQ_LEN * 2
    audio = port.slot(FS, -4400, iq, IQ_LEN, 0.2, af_size(120.0))["i16"]
    got = wc.decode(audio)
    assert [d["message"] for d in got] == ["K1ABC FN42 37"]
    assert abs(got[0]["freq_hz"] - 1500.0) < 0.4 and abs(got[0]["dt_s"]) < 0.1 and abs(got[0]["snr_db"] + 18.0) < 1.5
