"""Throughput of the EXACT kernel: 1 receiver x C channels, one FT8 slot resident in HBM."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cwsl_digi_b200 as cw
from cwsl_digi_b200 import synth

FS, IQ_LEN = 192000, 2048
nblk = 15 * FS // IQ_LEN
x = (torch.randn(nblk * IQ_LEN * 2, device="cuda") * 300).contiguous()
for C_ in (64, 256):
    rx = cw.Receiver(0, FS, IQ_LEN, mode=cw.MODE_EXACT)
    g = rx.add_group(15.0)
    for f in synth.stress_demod_freqs(C_):
        rx.add_channel(g, int(f), 0.9)
    rx.enable_timing(True)
    for it in range(2):
        rx.bind_device_iq(x.data_ptr(), nblk)
        rx.end_slot(g, None)
        rx.synchronize()
        t = rx.kernel_times()
        chs = nblk * IQ_LEN * C_
        print(f"exact C={C_} it={it}: demod {t['demod_ms']:.2f} ms -> {chs / t['demod_ms'] / 1e6:.1f} G ch-samples/s "
              f"= {15.0 / (t['demod_ms'] * 1e-3) * C_ / 1024:.0f} x real time per 1024 channels", flush=True)
    rx.close()
