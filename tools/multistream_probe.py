"""Reproduce/diagnose: many receivers, each on its private stream, 1024 channels, resident IQ."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cwsl_digi_b200 as cw
from cwsl_digi_b200 import synth

NRX = int(os.environ.get("NRX", "16"))
PRECOMMIT = os.environ.get("PRECOMMIT", "0") == "1"
MODE = cw.MODE_EXACT if os.environ.get("MODE", "fast") == "exact" else cw.MODE_FAST
SYNC_EACH = os.environ.get("SYNC_EACH", "0") == "1"
FS, IQ_LEN = 192000, 2048
nblk = 15 * FS // IQ_LEN
xs = [(torch.randn(nblk * IQ_LEN * 2, device="cuda") * 300).contiguous() for _ in range(NRX)]
torch.cuda.synchronize()
freqs = synth.stress_demod_freqs(1024)
rxs = []
for _ in range(NRX):
    rx = cw.Receiver(0, FS, IQ_LEN, mode=MODE)
    g = rx.add_group(15.0)
    for f in freqs:
        rx.add_channel(g, int(f), 0.9)
    rxs.append(rx)
if PRECOMMIT:
    for rx in rxs:
        rx.process(-1)      # forces commit (allocations, tables) before any kernel is in flight
        rx.synchronize()
for step in range(3):
    for rx, x in zip(rxs, xs):
        rx.bind_device_iq(x.data_ptr(), nblk)
        rx.end_slot(0, None)
        if SYNC_EACH:
            rx.synchronize()
    for rx in rxs:
        rx.synchronize()
    print("step", step, "ok", flush=True)
print("done")
