"""Small target for ncu: one receiver x C channels, one FT8 slot resident in HBM, a few launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cwsl_digi_b200 as cw
from cwsl_digi_b200 import synth

C_ = int(os.environ.get("PROF_CHANNELS", "1024"))
N_ = int(os.environ.get("PROF_SLOTS", "3"))
MODE = {"exact": cw.MODE_EXACT, "fast": cw.MODE_FAST, "stft": cw.MODE_STFT}[os.environ.get("PROF_MODE", "fast")]
FS, IQ_LEN = 192000, 2048
nblk = 15 * FS // IQ_LEN
x = (torch.randn(nblk * IQ_LEN * 2, device="cuda") * 300).contiguous()
rx = cw.Receiver(0, FS, IQ_LEN, mode=MODE)
g = rx.add_group(15.0)
for f in synth.stress_demod_freqs(C_):
    rx.add_channel(g, int(f), 0.9)
for _ in range(N_):
    rx.bind_device_iq(x.data_ptr(), nblk)
    rx.end_slot(g, None)
    rx.synchronize()
rx.close()
print("done")
