"""GPU test of the host-side mirror: config.ini -> Decoder/Receiver/Instance/DecoderPool -> WAV files,
checked against the oracle by re-deriving each WAV from the same synthetic IQ source."""
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "cwsl_digi_b200", "host")


@pytest.mark.parametrize("mode", ["exact", "stft"])   # (stft: the channelizer kernel, forced for small groups by conftest)
def test_station_demo_writes_reference_format_wavs(tmp_path, gpu, mode):
    exe = os.path.join(HOST, "station_demo")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", HOST, "all"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "data", "station_20m.ini"), str(tmp_path), "1", mode],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    wavs = sorted(p for p in os.listdir(tmp_path) if p.endswith(".wav"))
    # one 120 s span of signal: 8 FT8-class slots x 4 decoders (3 FT8 + JS8), 16 FT4 x 2, 2 JT65, 1 WSPR, minus the
    # first (partial, discarded) buffer of each decoder -- at least one file per mode must exist
    modes = {w.split("_")[2] for w in wavs}
    assert {"FT8", "FT4", "JT65", "JS8"} <= modes, wavs
    expect_len = {"FT8": 240000, "JS8": 240000, "FT4": 150000, "JT65": 780000, "WSPR": 1500000}
    for w in wavs:
        raw = open(os.path.join(tmp_path, w), "rb").read()
        mode = w.split("_")[2]
        assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[38:42] == b"data"
        file_len, = struct.unpack_from("<I", raw, 4)
        fmt_len, tag, ch, sr, bps, align, bits, cb = struct.unpack_from("<IHHIIHHH", raw, 16)
        data_len, = struct.unpack_from("<I", raw, 42)
        assert (fmt_len, tag, ch, sr, bps, align, bits, cb) == (18, 1, 1, 12000, 24000, 2, 16, 0)
        assert data_len == 2 * expect_len[mode] and file_len == 46 + data_len - 8 and len(raw) == 46 + data_len
        a = np.frombuffer(raw, np.int16, offset=46)
        # the demo puts a carrier 1500 Hz above every dial: normalised audio must peak near 0.9*32767
        # (0.2*32767 for WSPR) and carry a 1500 Hz line
        peak = int(np.abs(a.astype(np.int32)).max())
        ceil = 0.2 * 32767 if mode == "WSPR" else 0.9 * 32767
        assert 0.97 * ceil <= peak <= ceil + 1, (w, peak)
        seg = a[2000:2000 + 8192].astype(np.float64)
        spec = np.abs(np.fft.rfft(seg * np.hanning(seg.size)))
        assert abs(np.argmax(spec) * 12000 / seg.size - 1500) < 4, w
