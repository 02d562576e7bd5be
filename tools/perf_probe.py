"""Quick on-GPU probe: FP32 peak, exact/fast parity against the oracle, fast-kernel throughput."""
import json
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cwsl_digi_b200 as cw
from cwsl_digi_b200 import synth
from oracle.oracle import Ref, af_size

FS, IQ_LEN = 192000, 2048
res = {}
res["fp32_peak"] = cw.measure_fp32_peak(0)
print("fp32 peak", res["fp32_peak"], flush=True)

ref = Ref()
secs = float(os.environ.get("PROBE_SECS", "3"))
n = int(secs * FS) // IQ_LEN * IQ_LEN
freqs = [-26000, -4400, 90000, -96000]
iq = synth.receiver_iq(n, FS, freqs, receiver=0)
afs = af_size(15)
for mode, name in ((cw.MODE_EXACT, "exact"), (cw.MODE_FAST, "fast")):
    rx = cw.Receiver(0, FS, IQ_LEN, mode=mode)
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
        resid_db = 20 * np.log10(max(np.sqrt(np.mean(err ** 2)), 1e-30) / sig)
        mx, fac = rx.channel_stats(g, c)
        print(f"{name} f={f}: wi={wi}/{o['write_index']} bit-equal-raw={np.array_equal(raw.view(np.uint32), o['raw'].view(np.uint32))} "
              f"i16 maxdiff={d16.max()} ndiff={(d16 > 0).sum()} resid={resid_db:.1f} dB max={mx}/{o['max']} fac={fac}/{o['factor']}", flush=True)
        res[f"{name}_{f}"] = dict(maxdiff=int(d16.max()), ndiff=int((d16 > 0).sum()), resid_db=float(resid_db))
    rx.close()

# throughput: 1 receiver x C channels, one FT8 slot resident in HBM
for C_ in (64, 1024):
    freqs = synth.stress_demod_freqs(C_)
    nblk = 15 * FS // IQ_LEN
    x = (torch.randn(nblk * IQ_LEN * 2, device="cuda") * 300).contiguous()
    rx = cw.Receiver(0, FS, IQ_LEN, mode=cw.MODE_FAST)
    g = rx.add_group(15.0)
    for f in freqs:
        rx.add_channel(g, int(f), 0.9)
    rx.enable_timing(True)
    for it in range(3):
        rx.bind_device_iq(x.data_ptr(), nblk)
        rx.end_slot(g, None)
        rx.synchronize()
        t = rx.kernel_times()
        chs = nblk * IQ_LEN * C_
        print(f"C={C_} it={it} demod {t['demod_ms']:.3f} ms quant {t['quant_ms']:.3f} ms -> {chs / t['demod_ms'] / 1e6:.1f} G ch-samples/s (demod only), "
              f"{chs / (t['demod_ms'] + t['quant_ms']) / 1e6:.1f} incl quant", flush=True)
        res[f"fast_C{C_}"] = dict(demod_ms=t["demod_ms"], quant_ms=t["quant_ms"], gchs=chs / t["demod_ms"] / 1e6)
    rx.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/perf_probe.json", "w"), indent=1)
