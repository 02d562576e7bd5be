"""Generate tests/golden/*.npz from the reference's own headers (oracle/_ref/libcwsl_ref.so).

Run in the build container (needs /root/reference only to (re)build oracle/_ref):
    python tests/golden/make_golden.py
The vectors pin (a) the oracle restatement and (b) the CUDA path to the reference at commit time;
they hold the input IQ as well, so they do not depend on the synthetic generator staying stable.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cwsl_digi_b200 import synth  # noqa: E402
from oracle.oracle import Ref, af_size  # noqa: E402

CASES = [
    # name, fs, iq_len, period_s, n_iq_blocks, [(demod_freq, scale)], tones/ch
    ("ft8_192k", 192000, 2048, 15.0, 24, [(-26000, 0.90), (-10000, 0.90)], 4),
    ("wspr_192k_edges", 192000, 1024, 120.0, 40, [(-4400, 0.20), (-96000, 0.20), (90000, 0.20)], 2),
    ("ft4_96k", 96000, 1024, 7.5, 40, [(-20000, 0.90), (42000, 0.90)], 3),
    ("ft8_48k", 48000, 512, 15.0, 60, [(1500, 0.90), (-24000, 0.90)], 3),
]


def main():
    ref = Ref()
    for name, fs, iq_len, period, nblk, chans, ntones in CASES:
        n = nblk * iq_len
        freqs = [f for f, _ in chans]
        iq = synth.receiver_iq(n, fs, freqs, receiver=len(name), tones_per_channel=ntones)
        afs = af_size(period)
        raws, i16s, meta = [], [], []
        for f, sc in chans:
            o = ref.slot(fs, f, iq, iq_len, sc, afs)
            wi = o["write_index"]
            raws.append(o["raw"][:wi].copy())
            i16s.append(o["i16"][:wi].copy())
            assert not o["i16"][wi:].any()
            meta.append((f, sc, wi, o["max"], o["factor"]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), iq=iq, fs=fs, iq_len=iq_len, period=period,
                            af_size=afs, freqs=np.array([m[0] for m in meta], np.int32),
                            scales=np.array([m[1] for m in meta], np.float32),
                            write_index=np.array([m[2] for m in meta], np.int64),
                            maxval=np.array([m[3] for m in meta], np.float32),
                            factor=np.array([m[4] for m in meta], np.float32),
                            raw=np.stack(raws), i16=np.stack(i16s))
        print(name, "written:", n, "IQ samples,", len(chans), "channels, write_index", meta[0][2])


if __name__ == "__main__":
    main()
