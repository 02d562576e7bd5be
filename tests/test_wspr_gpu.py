"""north_star's decode gate, executable for WSPR: "the decode set for the synthetic WSPR slots must be identical to the
reference's".

A 120 s slot (1.44 M steps of the drifting float NCO recurrence) carries VALID WSPR transmissions -- real codewords
from the validated encoder in tests/wspr_codec.py -- from -8 dB down to -34 dB (in 2500 Hz; the decoder's threshold is
about -28 dB, wsprd's -29 ... -31 dB), at different frequencies and time lags, on two decoder channels. The slot goes
through the reference chain (oracle/_ref: the reference's own SSBD.hpp / LowPass.hpp + the restated prepareAudio /
int16 conversion) and through every GPU mode; both int16 buffers are handed to the same blind wsprd-style decoder
(candidate search, sync search, soft symbols, sequential decoding of the K = 32 code, unpacking). The DECODE SETS --
{(message, frequency to 0.01 Hz, lag to 1 ms)} -- must be equal, signal for signal, including which of the marginal
transmissions decode and which do not. The second variant adds an S9+40-like carrier 70 dB over the noise elsewhere in
the band (in STFT mode the dynamic-range guard then hands the quiet channels' segments to the direct-form kernel).

wsprd itself is not in the image (nor is anything of WSJT-X); FT8 / FT4 stay with the modulation-level proxy of
tests/test_decode_proxy.py because their LDPC generator cannot be derived offline."""
import numpy as np
import pytest

import wspr_codec as wc
from cwsl_digi_b200 import synth
from oracle.oracle import af_size

pytestmark = pytest.mark.gpu
FS, IQ_LEN, SIGMA, PERIOD = 192000, 2048, 300.0, 120.0
DEMOD = [-4400, 40000]                     # 14095600 WSPR under LO 14100000 (config.ini:77), and a second decoder
# (channel, call, grid, dBm, SNR dB in 2500 Hz, audio Hz of the signal's centre, start in the slot s)
PLAN = [
    (0, "K1ABC", "FN42", 37, -8.0, 1500.0, 1.0), (0, "W1AW", "FN31", 30, -15.0, 1432.1, 1.4),
    (0, "DL1ABC", "JO62", 23, -20.0, 1560.5, 0.6), (0, "G4JNT", "IO90", 10, -23.0, 1475.3, 1.2),
    (0, "VK3XYZ", "QF22", 0, -25.0, 1525.7, 2.0), (0, "JA1AA", "PM95", 60, -26.5, 1590.2, 0.9),
    (0, "N0ABC", "EM10", 33, -28.0, 1410.4, 1.1), (0, "F5ABC", "JN18", 27, -29.5, 1545.9, 1.3),
    (0, "9A1A", "JN75", 43, -31.0, 1452.6, 0.7), (0, "ZL1", "RF72", 7, -34.0, 1580.8, 1.5),
    (1, "EA8BFK", "IL38", 37, -12.0, 1510.0, 0.9), (1, "OH2XYZ", "KP20", 20, -24.0, 1447.3, 1.6),
    (1, "PY2AA", "GG66", 40, -27.0, 1571.9, 0.5), (1, "VE3ABC", "FN03", 13, -30.0, 1489.5, 1.2),
]
MUST_DECODE_DB = -25.0                     # every transmission at or above this must come back (decoder sanity)


@pytest.fixture(scope="module", params=[False, True], ids=["plain", "with_strong_carrier"])
def slot(request, ref):
    strong = request.param
    n = int(PERIOD * FS) // IQ_LEN * IQ_LEN
    x = synth.gaussian_iq(n, receiver=77, sigma=SIGMA)
    z = x[:, 0] + 1j * x[:, 1]
    del x
    for ch, call, grid, dbm, snr, fa, t0 in PLAN:
        ph, on = wc.fsk_audio_phase(wc.encode(call, grid, dbm), FS, DEMOD[ch] + fa - 1.5 * wc.TONE_HZ, t0, n)
        z[on] += np.sqrt(2.0 * SIGMA ** 2 * 2500.0 / FS * 10.0 ** (snr / 10.0)) * np.exp(1j * ph[on])
    if strong:
        t = np.arange(n, dtype=np.float64)
        z += np.sqrt(2.0 * SIGMA ** 2 * 2500.0 / FS * 1e7) * np.exp(2j * np.pi * ((-61000.0 * t) % FS) / FS)
    iq = np.ascontiguousarray(np.stack([z.real, z.imag], axis=1), np.float32).reshape(-1)
    del z
    want = [ref.slot(FS, f, iq, IQ_LEN, 0.2, af_size(PERIOD))["i16"] for f in DEMOD]
    sets = [wc.decode_set(a) for a in want]
    return dict(strong=strong, iq=iq, want=want, sets=sets)


def test_decoder_sees_the_transmissions(slot):
    """The gate is only worth something if the decoder really decodes: everything at or above MUST_DECODE_DB comes
    back from the reference chain's audio, nothing that was not sent does, and the weakest transmissions do not."""
    sent = {ch: {f"{c} {g} {p}": snr for k, c, g, p, snr, *_ in PLAN if k == ch} for ch in range(len(DEMOD))}
    for ch, s in enumerate(slot["sets"]):
        got = {m for m, _, _ in s}
        assert got <= set(sent[ch]), got - set(sent[ch])                   # no false decodes
        assert {m for m, snr in sent[ch].items() if snr >= MUST_DECODE_DB} <= got
        print(f"channel {ch} (strong carrier {slot['strong']}): decoded {sorted(sent[ch][m] for m in got)} dB, "
              f"missed {sorted(snr for m, snr in sent[ch].items() if m not in got)} dB")
    assert any(f"{c} {g} {p}" not in {m for m, _, _ in slot["sets"][k]} for k, c, g, p, *_ in PLAN)


@pytest.mark.parametrize("mode", ["exact", "fast", "stft"])
def test_wspr_decode_set_identical_to_reference(gpu, slot, mode):
    cw = gpu
    m = {"exact": cw.MODE_EXACT, "fast": cw.MODE_FAST, "stft": cw.MODE_STFT}[mode]
    with cw.Receiver(0, FS, IQ_LEN, mode=m) as rx:
        g = rx.add_group(PERIOD)
        for f in DEMOD:
            rx.add_channel(g, f, 0.2)                                      # wspraudioscalefactor, source/Instance.cpp:320-329
        rx.push_iq(slot["iq"])
        out, wi = rx.end_slot_numpy(g)
        redone = rx.guard_stats(g)["redone"] if mode == "stft" else 0
    if mode == "stft" and slot["strong"]:
        assert redone > 0
    report = []
    for ch in range(len(DEMOD)):
        if mode == "exact":
            assert np.array_equal(out[ch], slot["want"][ch])
        d = np.abs(out[ch].astype(np.int32) - slot["want"][ch].astype(np.int32))
        assert d.max() <= 1
        got = wc.decode_set(out[ch])
        report.append((ch, int((d > 0).sum()), len(got), sorted(got ^ slot["sets"][ch])))
    print(f"WSPR {mode}, strong carrier {slot['strong']}: (channel, int16 samples that differ, decodes, set difference) {report}")
    for ch, _, _, diff in report:
        assert not diff, (ch, diff)
