"""Small target for compute-sanitizer (memcheck / racecheck / synccheck): all three demodulator kernels, several
tiles and segments, channel groups, ring wrap, partial last tile / batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CWSL_STFT_MIN_CHANNELS", "1")
import numpy as np

import cwsl_digi_b200 as cw
from cwsl_digi_b200 import synth

FS, IQ_LEN = 192000, 2048
freqs = [int(f) for f in synth.stress_demod_freqs(40)]          # 2 channel groups (32 + 8)
iq = synth.receiver_iq(70 * IQ_LEN, FS, freqs[:2], receiver=0, tones_per_channel=1)
for mode in (cw.MODE_FAST, cw.MODE_EXACT, cw.MODE_STFT):
    with cw.Receiver(0, FS, IQ_LEN, ring_seconds=0.3, mode=mode) as rx:
        g = rx.add_group(15.0)
        for f in freqs:
            rx.add_channel(g, f, 0.9)
        for b in range(0, 70, 7):
            rx.push_iq(iq[b * IQ_LEN * 2:(b + 7) * IQ_LEN * 2])
        out, wi = rx.end_slot_numpy(g)
        print("mode", mode, "wi", wi, "checksum", int(out.astype(np.int64).sum()))
        # second slot on the same handle through the packed hand-off (rows of write_index samples, 1-D copy)
        for b in range(0, 42, 7):
            rx.push_iq(iq[b * IQ_LEN * 2:(b + 7) * IQ_LEN * 2])
        packed = np.zeros(len(freqs) * rx.group_af_size(g), np.int16)
        wi = rx.end_slot_packed(g, packed)
        rx.synchronize()
        print("mode", mode, "packed wi", wi, "checksum", int(packed[:len(freqs) * wi].astype(np.int64).sum()))
# the STFT guard's redo path: one carrier ~80 dB over the noise, so the quiet channels' segments are recomputed by the
# indirect FAST launch (work lists built on the device, phases replayed from the anchors into the per-stream scratch)
hdr = iq.copy()
t = np.arange(hdr.size // 2, dtype=np.float64)
ph = 2 * np.pi * ((freqs[5] + 1500.0) * t % FS) / FS
hdr[0::2] = hdr[0::2] * 0.01 + (3.0e4 * np.cos(ph)).astype(np.float32)
hdr[1::2] = hdr[1::2] * 0.01 + (3.0e4 * np.sin(ph)).astype(np.float32)
with cw.Receiver(0, FS, IQ_LEN, ring_seconds=0.3, mode=cw.MODE_STFT) as rx:
    g = rx.add_group(15.0)
    for f in freqs:
        rx.add_channel(g, f, 0.9)
    for b in range(0, 70, 7):
        rx.push_iq(hdr[b * IQ_LEN * 2:(b + 7) * IQ_LEN * 2])
    out, wi = rx.end_slot_numpy(g)
    print("stft guard redo: wi", wi, "checksum", int(out.astype(np.int64).sum()), "guard", rx.guard_stats(g))
# the channelizer's other geometries (two / four hops per FFT warp): 96 and 48 kHz receivers
for fs, il in ((96000, 1024), (48000, 512)):
    fr = [int(f) for f in np.linspace(-fs // 2, fs // 2 - 6000, 20)]
    x = synth.receiver_iq(70 * il, fs, fr[:2], receiver=1, tones_per_channel=1)
    with cw.Receiver(0, fs, il, ring_seconds=0.3, mode=cw.MODE_STFT) as rx:
        g = rx.add_group(15.0)
        for f in fr:
            rx.add_channel(g, f, 0.9)
        for b in range(0, 70, 7):
            rx.push_iq(x[b * il * 2:(b + 7) * il * 2])
        out, wi = rx.end_slot_numpy(g)
        print("stft fs", fs, "wi", wi, "checksum", int(out.astype(np.int64).sum()))
print("done")
