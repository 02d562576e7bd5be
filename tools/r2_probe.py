"""Round-2 GPU probe (one receiver, 1024 channels, one FT8 slot resident in HBM):
  (1) error floor of the raw STFT channelizer (guard off) and of the FAST kernel against the bit-exact EXACT mode,
      per channel, as a fraction of the band's rms -- for several input classes; this is what the guard threshold
      (cwsl_guard.cu, include/cwsl_b200.h) is derived from;
  (2) what the guard does on each input (channel segments redone) and the residual of the guarded STFT mode;
  (3) kernel times per mode (CUDA events inside the library).
Writes gpurun_out/r2_probe.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cwsl_digi_b200 as cw
from cwsl_digi_b200 import synth

FS, IQ_LEN = 192000, 2048
NBLK = 15 * FS // IQ_LEN
N = NBLK * IQ_LEN
C_ = int(os.environ.get("PROBE_CHANNELS", "1024"))
freqs = synth.stress_demod_freqs(C_)
t = torch.arange(N, device="cuda", dtype=torch.float64)


def tone(x, hz, amp, gate=None):
    ph = 2 * np.pi * ((hz * t) % FS) / FS
    a = amp if gate is None else amp * gate
    x[0::2] += (a * torch.cos(ph)).float()
    x[1::2] += (a * torch.sin(ph)).float()


def make(kind):
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn(2 * N, device="cuda", generator=g) * 300.0
    if kind == "stress16":                       # tests/test_parity_gpu.py::stress_run
        for j, f in enumerate(freqs[::64]):
            tone(x, int(f) + 700 + 100 * j, 4000.0)
    elif kind == "bench":                        # bench.py: 8 tones of amplitude 8000 in every passband
        rng = np.random.default_rng(synth.BASE_SEED)
        hz = (freqs[:, None].astype(np.float64) + rng.uniform(200.0, 2900.0, (C_, 8))).reshape(-1)
        bins = torch.from_numpy(np.mod(np.rint(hz * N / FS).astype(np.int64), N)).cuda()
        ph = torch.rand(bins.numel(), device="cuda", generator=g, dtype=torch.float64) * (2 * np.pi)
        spec = torch.zeros(N, device="cuda", dtype=torch.complex64)
        spec.index_add_(0, bins, (8000.0 * torch.exp(1j * ph)).to(torch.complex64))
        z = torch.fft.ifft(spec, norm="forward")
        x[0::2] += z.real
        x[1::2] += z.imag
    elif kind == "carrier80":                    # one carrier 80 dB above the noise in a 6 kHz channel
        tone(x, int(freqs[C_ // 3]) + 1500, 300.0 * np.sqrt(2 * 6000 / FS) * 1e4)
    elif kind == "noise":
        pass
    elif kind == "two_strong":                   # two S9+60-like carriers, everything else quiet
        tone(x, 50000 + 333, 3.0e5)
        tone(x, -70000 + 777, 1.0e5)
    return x.contiguous()


def run(x, mode, guard_db=None, want_raw=True):
    with cw.Receiver(0, FS, IQ_LEN, mode=mode) as rx:
        grp = rx.add_group(15.0)
        for f in freqs:
            rx.add_channel(grp, int(f), 0.9)
        if guard_db is not None:
            rx.set_stft_guard(guard_db)
        rx.enable_timing(True)
        for _ in range(2):
            rx.bind_device_iq(x.data_ptr(), NBLK)
            rx.end_slot(grp, None)
            rx.synchronize()
            kt = rx.kernel_times()
        st = rx.guard_stats(grp) if mode == cw.MODE_STFT else None
        q = torch.empty((C_, rx.group_af_size(grp)), dtype=torch.int16, device="cuda")
        for c in range(C_):
            rx.copy_device_audio(grp, c, q[c].data_ptr())
        rx.synchronize()
        raw = np.stack([rx.read_float_audio(grp, c) for c in range(C_)]) if want_raw else None
    return q.cpu().numpy(), raw, kt, st


out = {}
wi = N // 16
for kind in os.environ.get("PROBE_KINDS", "stress16,bench,carrier80,noise,two_strong").split(","):
    x = make(kind)
    band_rms = float(torch.sqrt((x.double() ** 2).mean() * 2).item())          # rms of |x|
    q_e, r_e, kt_e, _ = run(x, cw.MODE_EXACT)
    res = dict(band_rms=band_rms, exact_ms=kt_e)
    sig = np.sqrt((r_e[:, :wi].astype(np.float64) ** 2).mean(axis=1))
    res["channel_rms_over_band_db"] = dict(min=float(20 * np.log10(sig.min() / band_rms)),
                                           median=float(20 * np.log10(np.median(sig) / band_rms)),
                                           max=float(20 * np.log10(sig.max() / band_rms)))
    for name, mode, gdb in (("fast", cw.MODE_FAST, None), ("stft_raw", cw.MODE_STFT, 0.0), ("stft_guard", cw.MODE_STFT, None)):
        q, r, kt, st = run(x, mode, gdb)
        err = np.sqrt(((r[:, :wi].astype(np.float64) - r_e[:, :wi]) ** 2).mean(axis=1))
        resid = 20 * np.log10(np.maximum(err, 1e-300) / sig)
        d = np.abs(q.astype(np.int32) - q_e.astype(np.int32))
        res[name] = dict(kernel_ms=kt, guard=st,
                         err_over_band_rms=dict(max=float((err / band_rms).max()), median=float(np.median(err / band_rms)),
                                                p99=float(np.quantile(err / band_rms, 0.99))),
                         resid_db=dict(worst=float(resid.max()), median=float(np.median(resid))),
                         worst_channel=int(resid.argmax()),
                         max_lsb=int(d.max()), frac_diff=float((d > 0).mean()))
        print(kind, name, json.dumps(res[name]), flush=True)
    out[kind] = res
    del x
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r2_probe.json", "w"), indent=1)
print("probe done")
