"""On-GPU probe of the STFT channelizer: parity of a few channels against the reference chain, then throughput of
1 receiver x C channels x one FT8 slot resident in HBM (demod / quantise kernel times from CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CWSL_STFT_MIN_CHANNELS", "1")
import numpy as np
import torch

import cwsl_digi_b200 as cw
from cwsl_digi_b200 import synth
from oracle.oracle import Ref, af_size

FS = int(os.environ.get("PROBE_FS", "192000"))
IQ_LEN = 2048 * FS // 192000
if os.environ.get("PROBE_PARITY", "1") == "1":
    ref = Ref()
    n = int(2 * FS) // IQ_LEN * IQ_LEN
    freqs = [-96000, -95818, -26000, -4400, 0, 37, 12345, 36000, 89636, 90000]
    iq = synth.receiver_iq(n, FS, freqs, receiver=0)
    afs = af_size(15)
    rx = cw.Receiver(0, FS, IQ_LEN, mode=cw.MODE_STFT)
    g = rx.add_group(15.0)
    for f in freqs:
        rx.add_channel(g, f, 0.9)
    rx.push_iq(iq)
    out, wi = rx.end_slot_numpy(g)
    for c, f in enumerate(freqs):
        o = ref.slot(FS, f, iq, IQ_LEN, 0.9, afs)
        raw = rx.read_float_audio(g, c)
        d16 = np.abs(out[c].astype(np.int32) - o["i16"].astype(np.int32))
        err = raw.astype(np.float64) - o["raw"].astype(np.float64)
        sig = np.sqrt(np.mean(o["raw"].astype(np.float64) ** 2))
        r = 20 * np.log10(max(np.sqrt(np.mean(err ** 2)), 1e-30) / sig)
        mx, fac = rx.channel_stats(g, c)
        print(f"stft f={f}: wi={wi}/{o['write_index']} i16 maxdiff={d16.max()} ndiff={(d16 > 0).sum()} resid={r:.1f} dB "
              f"max={mx}/{o['max']}", flush=True)
    rx.close()

for C_ in [int(v) for v in os.environ.get("PROBE_CHANNELS", "64,256,1024,4096").split(",")]:
    freqs = np.rint(np.linspace(-FS // 2, FS // 2 - 6000, C_)).astype(np.int64) if FS != 192000 else synth.stress_demod_freqs(C_)
    nblk = 15 * FS // IQ_LEN
    x = (torch.randn(nblk * IQ_LEN * 2, device="cuda") * 300).contiguous()
    for mode, name in ((cw.MODE_FAST, "fast"), (cw.MODE_STFT, "stft")):
        if name == "fast" and C_ > 1024:
            continue
        rx = cw.Receiver(0, FS, IQ_LEN, mode=mode)
        g = rx.add_group(15.0)
        for f in freqs:
            rx.add_channel(g, int(f), 0.9)
        rx.enable_timing(True)
        best = None
        for it in range(4):
            rx.bind_device_iq(x.data_ptr(), nblk)
            rx.end_slot(g, None)
            rx.synchronize()
            t = rx.kernel_times()
            t = (t['demod_ms'], t['quant_ms'])
            if it and (best is None or t[0] < best[0]):
                best = t
        chs = C_ * nblk * IQ_LEN
        print(f"{name} C={C_}: demod {best[0]:.3f} ms quantise {best[1]:.3f} ms -> {chs / best[0] / 1e6:.1f} G ch-samples/s "
              f"(demod only), {chs / (best[0] + best[1]) / 1e6:.1f} G with quantise = "
              f"{15.0 / ((best[0] + best[1]) * 1e-3) * C_ / 1024:.0f} x real time per 1024 channels", flush=True)
        rx.close()
