"""GPU tests of the host-side mirror: config.ini -> Decoder / Receiver / Instance /
DecoderPool -> the artefacts the external decoders are handed. The station demo dumps the IQ every receiver was fed
and the IQ-block index of every slot edge, so each artefact is re-derived from the same samples with the oracle:
  f1  WAV files: 46-byte reference header + payload == the oracle's int16 vector, sample for sample (EXACT mode);
  f2  transfermethod=shmem: what the jt9 stand-in finds in d2[] of the shared-memory block == the oracle's int16,
      parameters per mode, ipc[] handshake completed for every item; WSPR / JS8 still arrive as WAV files;
  f3  the same with the IQ coming through the CWSL shared-memory ring (producer thread -> POSIX segment ->
      CwslShmSource -> pinned staging ring -> cwsl_rx_push_iq), no device synchronisation per block;
  f4  the WSPR decoders of the demo hear valid WSPR transmissions (WsprSynth.hpp): the WAV files wsprd would be
      given decode -- with the blind decoder of tests/wspr_codec.py -- to exactly the transmitted messages, and to the
      same decode set in all three arithmetic modes."""
import glob
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import wspr_codec as wc
from oracle.oracle import af_size

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "cwsl_digi_b200", "host")
FS, IQ_LEN = 192000, 2048
PERIOD = {"FT8": 15.0, "JS8": 15.0, "FT4": 7.5, "JT65": 60.0, "WSPR": 120.0}
EXPECT_LEN = {"FT8": 240000, "JS8": 240000, "FT4": 150000, "JT65": 780000, "WSPR": 1500000}


def run_demo(tmp_path, *args, ini="station_20m.ini"):
    exe = os.path.join(HOST, "station_demo")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", HOST, "all"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "data", ini), str(tmp_path), *args],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def read_wav(path):
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[38:42] == b"data"
    file_len, = struct.unpack_from("<I", raw, 4)
    fmt = struct.unpack_from("<IHHIIHHH", raw, 16)
    data_len, = struct.unpack_from("<I", raw, 42)
    assert fmt == (18, 1, 1, 12000, 24000, 2, 16, 0)             # source/WaveFile.hpp:19-35
    assert file_len == 46 + data_len - 8 and len(raw) == 46 + data_len
    return np.frombuffer(raw, np.int16, offset=46)


def expected_slots(tmp_path, ref, decoders):
    """{(mode, dial): [int16 vector per finished slot]} from the dumped IQ and slot edges of every receiver.
    decoders: [(dial_hz, mode, scale)]. The first edge of a group ends the partial start-up buffer, which the
    reference discards (source/Instance.cpp:224-227)."""
    out = {}
    for iq_path in glob.glob(os.path.join(tmp_path, "iq_*.f32")):
        lo = int(re.search(r"iq_(\d+)\.f32", iq_path).group(1))
        iq = np.fromfile(iq_path, np.float32)
        edges = {}
        for line in open(os.path.join(tmp_path, f"edges_{lo}.txt")):
            per, blk = line.split()
            edges.setdefault(float(per), []).append(int(blk))
        n_blocks = iq.size // (2 * IQ_LEN)
        for dial, mode, scale in decoders:
            if (dial + 50000) // 100000 * 100000 != lo:
                continue
            e = [b for b in edges[PERIOD[mode]] if b < n_blocks]
            slots = []
            for b0, b1 in zip(e[:-1], e[1:]):                   # blocks b0+1 .. b1 inclusive
                span = iq[(b0 + 1) * IQ_LEN * 2:(b1 + 1) * IQ_LEN * 2]
                slots.append(ref.slot(FS, dial - lo, span, IQ_LEN, scale, af_size(PERIOD[mode]))["i16"])
            out[(mode, dial)] = slots
    return out


def iq_digest(tmp_path):
    import hashlib
    h = hashlib.md5()
    for f in sorted(glob.glob(os.path.join(tmp_path, "iq_*.f32")) + glob.glob(os.path.join(tmp_path, "edges_*.txt"))):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            for chunk in iter(lambda: fh.read(1 << 24), b""):
                h.update(chunk)
    return h.hexdigest()


@pytest.fixture(scope="module")
def station(tmp_path_factory, gpu, ref):
    """One EXACT-mode run of the demo (WAV hand-off) + the oracle's vectors for every slot of every decoder; the
    other runs feed the same seeded IQ (checked by digest), so they reuse the oracle's vectors."""
    d = tmp_path_factory.mktemp("station_exact")
    run_demo(d, "1", "exact", "wavefile", "dumpiq")
    return dict(dir=d, want=expected_slots(d, ref, DECODERS_20M), digest=iq_digest(d))


DECODERS_20M = [(14095600, "WSPR", 0.20), (14090000, "FT8", 0.90), (14080000, "FT4", 0.90), (14074000, "FT8", 0.90),
                (14076000, "JT65", 0.90), (14078000, "JS8", 0.90), (7074000, "FT8", 0.90), (7047500, "FT4", 0.90)]


# what station_demo transmits to every WSPR decoder (kWsprDemoTransmissions): message -> (audio Hz, SNR dB, lag s)
WSPR_DEMO = {"K1ABC FN42 37": (1500.0, -12.0, 0.0), "W1AW FN31 30": (1440.0, -20.0, 0.3), "G4JNT IO90 23": (1570.0, -24.0, -0.2)}


def wspr_decode_sets(tmp_path):
    """{wav name without the slot stamp: decode set} of the WSPR hand-off files of one demo run."""
    out = {}
    for w in sorted(glob.glob(os.path.join(tmp_path, "*_WSPR_*.wav"))):
        out.setdefault(os.path.basename(w).split("_")[1], []).append(wc.decode_set(read_wav(w)))
    return out


def match_all(got, want, what):
    """Every artefact equals exactly one expected slot of its decoder, no slot is used twice, none is missing."""
    for key, vecs in got.items():
        left = list(range(len(want[key])))
        for v in vecs:
            hit = [i for i in left if want[key][i].size == v.size and np.array_equal(want[key][i], v)]
            assert hit, f"{what} of {key}: payload matches no oracle slot"
            left.remove(hit[0])
    for key, slots in want.items():
        assert len(got.get(key, [])) == len(slots), (what, key, len(got.get(key, [])), len(slots))


def test_wav_payload_equals_oracle_int16(station):
    """f1 (source/DecoderPool.hpp:900-964): EXACT mode, every WAV the pool writes is the oracle's vector."""
    tmp_path, want = station["dir"], station["want"]
    got = {}
    for w in sorted(glob.glob(os.path.join(tmp_path, "*.wav"))):
        _, freq, mode, _, _ = os.path.basename(w)[:-4].split("_")
        a = read_wav(w)
        assert a.size == EXPECT_LEN[mode]
        got.setdefault((mode, int(freq)), []).append(a)
    assert {k[0] for k in got} == {"FT8", "FT4", "JT65", "JS8", "WSPR"}
    match_all(got, want, "WAV")


def test_wspr_wav_decodes_to_the_transmitted_messages(station):
    """f4 + f1: config.ini -> Receiver -> GPU front-end -> WAV for wsprd; the file decodes to the three valid
    transmissions the synthetic source put on the air, at their frequencies, lags and SNRs."""
    wavs = sorted(glob.glob(os.path.join(station["dir"], "*_WSPR_*.wav")))
    assert wavs
    for w in wavs:
        got = {d["message"]: d for d in wc.decode(read_wav(w))}
        assert set(got) == set(WSPR_DEMO), (w, sorted(got))
        for msg, (fa, snr, lag) in WSPR_DEMO.items():
            d = got[msg]
            # (the slot edge falls on an IQ-block boundary: the lag is known to one block = 10.7 ms)
            assert abs(d["freq_hz"] - fa) < 0.4 and abs(d["dt_s"] - lag) < 0.12 and abs(d["snr_db"] - snr) < 1.5, d


def test_jt9_shared_memory_handoff_equals_oracle(tmp_path, station):
    """f2 (source/DecoderPool.hpp:421-448, :575-593, :689-709): the jt9 stand-in attaches by key, finds the oracle's
    samples in d2[] with the mode's parameters, and every handshake completes; f3: the IQ came through the CWSL
    shared-memory ring and the receiver's pinned staging ring."""
    out = run_demo(tmp_path, "1", "exact", "shmem", "dumpiq", "shmsource")
    assert iq_digest(tmp_path) == station["digest"]              # the ring delivered every block, in order, unchanged
    want = station["want"]
    got, n_shm = {}, 0
    nmode = {"FT8": 8, "FT4": 5, "JT65": 65}
    for t in sorted(glob.glob(os.path.join(tmp_path, "CWSL_DIGI_*.txt"))):
        kv = dict(p.split("=") for p in open(t).read().split())
        mode, freq = kv["mode"], int(kv["freq"])
        assert kv["sane"] == "1" and int(kv["nmode"]) == nmode[mode] and int(kv["ntrperiod"]) == int(PERIOD[mode])
        assert int(kv["nfb"]) == 3000 and int(kv["ndepth"]) == 3
        d2 = np.fromfile(t[:-4] + ".d2", np.int16)
        assert d2.size == EXPECT_LEN[mode]
        got.setdefault((mode, freq), []).append(d2)
        n_shm += 1
    for w in sorted(glob.glob(os.path.join(tmp_path, "*.wav"))):
        _, freq, mode, _, _ = os.path.basename(w)[:-4].split("_")
        assert mode in ("WSPR", "JS8"), w                        # source/DecoderPool.hpp:379-395
        got.setdefault((mode, int(freq)), []).append(read_wav(w))
    match_all(got, want, "hand-off")
    m = re.search(r"(\d+) of them through jt9 shared memory \((\d+) handshakes completed\)", out)
    assert m and int(m.group(1)) == n_shm == int(m.group(2)) and n_shm > 0
    assert not glob.glob("/dev/shm/CWSL_DIGI_*") and not glob.glob("/dev/shm/CWSL*_demo*")   # nothing left behind


@pytest.mark.parametrize("mode", ["fast", "stft"])   # (stft: the channelizer kernel, forced for small groups by conftest)
def test_station_demo_fast_modes_within_one_lsb(tmp_path, station, mode):
    run_demo(tmp_path, "1", mode, "wavefile", "dumpiq")
    assert iq_digest(tmp_path) == station["digest"]
    want = station["want"]
    n = 0
    for w in sorted(glob.glob(os.path.join(tmp_path, "*.wav"))):
        _, freq, m, _, _ = os.path.basename(w)[:-4].split("_")
        a = read_wav(w).astype(np.int32)
        best = min(int(np.abs(a - s.astype(np.int32)).max()) for s in want[(m, int(freq))])
        assert best <= 1, (w, best)
        n += 1
    assert n == sum(len(v) for v in want.values())
    assert wspr_decode_sets(tmp_path) == wspr_decode_sets(station["dir"])      # the decode gate, through the host mirror
