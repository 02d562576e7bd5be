"""Where does a receiver's 0.69 ms in the 64-receiver step go? R receivers x 1024 channels share one high-priority
stream like bench.py; per variant (library events on/off, STFT guard on/off, quantise pass joined per receiver or not) the
wall-clock-free device time per receiver between two events on that stream."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cwsl_digi_b200 as cw
from cwsl_digi_b200 import synth

FS, IQ_LEN, R = 192000, 2048, int(os.environ.get("GAP_RX", "16"))
NBLK = 15 * FS // IQ_LEN
N = NBLK * IQ_LEN
freqs = synth.stress_demod_freqs(1024)
stream = torch.cuda.Stream(priority=-1)
torch.cuda.set_stream(stream)
xs = [(torch.randn(2 * N, device="cuda") * 300.0).contiguous() for _ in range(R)]


def run(label, timing, guard_db, join_each=False, steps=4, pipelined=False):
    rxs = []
    for _ in range(R):
        rx = cw.Receiver(0, FS, IQ_LEN, mode=cw.MODE_STFT)
        g = rx.add_group(15.0)
        for f in freqs:
            rx.add_channel(g, int(f), 0.9)
        rx.set_stream(stream.cuda_stream)
        if guard_db is not None:
            rx.set_stft_guard(guard_db)
        rxs.append(rx)

    def step():
        if pipelined:   # the bulk of receiver r+1's demodulation is queued BEFORE receiver r's quantise pass
            rxs[0].bind_device_iq(xs[0].data_ptr(), NBLK)
            rxs[0].process(0)
            for i in range(R):
                if i + 1 < R:
                    rxs[i + 1].bind_device_iq(xs[i + 1].data_ptr(), NBLK)
                    rxs[i + 1].process(0)
                rxs[i].end_slot(0, None)
            return
        for rx, x in zip(rxs, xs):
            rx.bind_device_iq(x.data_ptr(), NBLK)
            rx.end_slot(0, None)
            if join_each:
                rx.join_output()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    for rx in rxs:
        rx.enable_timing(timing)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    for rx in rxs:
        rx.join_output()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (steps * R)
    kt = None
    if timing:
        kts = [rx.kernel_times() for rx in rxs]
        kt = {k: sum(t[k] for t in kts) / (steps * R) for k in ("demod_ms", "main_ms", "guard_pre_ms", "guard_post_ms", "quant_ms")}
    for rx in rxs:
        rx.close() if hasattr(rx, "close") else None
    print(json.dumps(dict(variant=label, ms_per_receiver=round(ms, 4), library_events=kt)), flush=True)
    del rxs


run("pipelined (process r+1 before end_slot r), events on, guard on", True, None, pipelined=True)
run("pipelined, events off, guard off", False, 0.0, pipelined=True)
run("events on, guard on (bench)", True, None)
run("events off, guard on", False, None)
run("events off, guard off", False, 0.0)
run("events on, guard off", True, 0.0)
run("events off, guard on, quantise joined after every receiver (no overlap)", False, None, join_each=True)
