"""x real time of BASELINE.json configs[0] (1 FT8 decoder) and configs[1] (default 20 m set) on one GPU."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cwsl_digi_b200 as cw

FS, IQ_LEN = 192000, 2048
LO = 14100000


def run(name, decs, seconds, mode):
    nblk = int(seconds * FS) // IQ_LEN
    x = (torch.randn(nblk * IQ_LEN * 2, device="cuda") * 300).contiguous()
    h = torch.empty(nblk * IQ_LEN * 2, dtype=torch.float32).pin_memory()
    h.copy_(x)
    rx = cw.Receiver(0, FS, IQ_LEN, mode=mode)
    groups = {}
    for dial, per, sc in decs:
        if per not in groups:
            groups[per] = rx.add_group(per)
        rx.add_channel(groups[per], dial - LO, sc)
    outs = {per: cw.HostBuffer(rx.num_channels(g), rx.group_af_size(g)) for per, g in groups.items()}
    best = 1e9
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        # one hyper-period from HOST memory: push everything, end each group's slots at its own edges
        pos = 0
        step = int(7.5 * FS) // IQ_LEN
        k = 0
        while pos < nblk:
            nb = min(step, nblk - pos)
            rx.push_iq((h.data_ptr() + pos * IQ_LEN * 8, nb))
            pos += nb
            k += 1
            for per, g in groups.items():
                if (k * 7.5) % per == 0:
                    rx.end_slot(g, outs[per].ptr)
        rx.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{name} [{'exact' if mode == cw.MODE_EXACT else 'fast'}]: {seconds:.0f} s of signal, {len(decs)} decoders, "
          f"host->host in {best * 1e3:.2f} ms = {seconds / best:.0f} x real time", flush=True)
    rx.close()


for mode in (cw.MODE_FAST, cw.MODE_EXACT):
    run("config 1 (FT8 @ 14074000)", [(14074000, 15.0, 0.9)], 15.0, mode)
    run("config 2 (default 20 m set)", [(14095600, 120.0, 0.2), (14090000, 15.0, 0.9), (14080000, 7.5, 0.9), (14074000, 15.0, 0.9),
                                        (14076000, 60.0, 0.9), (14078000, 15.0, 0.9), (14097000, 120.0, 0.9)], 120.0, mode)
