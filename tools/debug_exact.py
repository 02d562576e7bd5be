import os, sys, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cwsl_digi_b200 as cw
for path in sorted(glob.glob("tests/golden/*.npz")):
    g = np.load(path)
    fs, iq_len = int(g["fs"]), int(g["iq_len"])
    with cw.Receiver(0, fs, iq_len, mode=cw.MODE_EXACT) as rx:
        grp = rx.add_group(float(g["period"]))
        for f, s in zip(g["freqs"], g["scales"]):
            rx.add_channel(grp, int(f), float(s))
        rx.push_iq(g["iq"])
        out, wi = rx.end_slot_numpy(grp)
        for c in range(len(g["freqs"])):
            raw = rx.read_float_audio(grp, c)[:wi]
            want = g["raw"][c]
            bad = np.nonzero(raw.view(np.uint32) != want.view(np.uint32))[0]
            print(os.path.basename(path), "ch", c, "wi", wi, "mismatches", bad.size, "first", bad[:8],
                  "maxabs", float(np.abs(raw - want).max()), flush=True)
            if bad.size:
                i = bad[0]
                print("   got", raw[i:i+4], "want", want[i:i+4])
