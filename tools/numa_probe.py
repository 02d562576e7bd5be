"""Host-side probe for the end-to-end leg at N > 1: concurrent pinned D2H/H2D bandwidth per rank with the hand-off
buffer allocated (a) wherever the process happens to run, (b) on the GPU's own NUMA node, (c) on another node.
Run under torchrun: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/numa_probe.py"""
import glob
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist


def cpus_of_node(node):
    txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = set()
    for part in txt.split(","):
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    buf = ctypes.create_string_buffer(32)
    rt.cudaDeviceGetPCIBusId(buf, 32, local)
    busid = buf.value.decode().lower()
    try:
        node = int(open(f"/sys/bus/pci/devices/{busid}/numa_node").read())
    except Exception as e:  # noqa: BLE001
        node = -1
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    if rank == 0:
        print("nodes", nodes, "cpus", os.cpu_count(), flush=True)
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)
    all_cpus = os.sched_getaffinity(0)
    print(f"rank {rank} gpu {local} bus {busid} numa_node {node} affinity {len(all_cpus)} cpus", flush=True)

    n = 1 << 30
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")

    results = {}

    def run(tag, cpus):
        if cpus:
            os.sched_setaffinity(0, cpus)
        host = torch.empty(n, dtype=torch.uint8).pin_memory()
        host.zero_()
        os.sched_setaffinity(0, all_cpus)
        res = {}
        for name, (dst, src) in {"d2h": (host, dev), "h2d": (dev, host)}.items():
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(10):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            res[name] = 10 * n / dt / 1e9
            if world > 1:
                dist.barrier()
        # both directions at once (the end-to-end leg pushes the next receiver's IQ while a slot is copied back)
        host2 = torch.empty(n // 16, dtype=torch.uint8).pin_memory()
        dev2 = torch.empty(n // 16, dtype=torch.uint8, device="cuda")
        s2 = torch.cuda.Stream()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            host.copy_(dev, non_blocking=True)
            with torch.cuda.stream(s2):
                dev2.copy_(host2, non_blocking=True)
        torch.cuda.synchronize()
        res["d2h_duplex"] = 10 * n / (time.perf_counter() - t0) / 1e9
        if world > 1:
            dist.barrier()
        print(f"rank {rank} {tag}: d2h {res['d2h']:.1f} GB/s  h2d {res['h2d']:.1f} GB/s  d2h with 1/16 h2d beside it "
              f"{res['d2h_duplex']:.1f} GB/s", flush=True)
        results[tag] = res
        del host, host2, dev2

    run("default placement", None)
    if node >= 0 and len(nodes) > 1:
        run(f"local node {node}", cpus_of_node(node) & all_cpus or None)
        other = [x for x in nodes if x != node][0]
        run(f"remote node {other}", cpus_of_node(other) & all_cpus or None)
    # solo: one rank at a time
    for r in range(world):
        if world > 1:
            dist.barrier()
        if r == rank:
            host = torch.empty(n, dtype=torch.uint8).pin_memory()
            host.copy_(dev, non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                host.copy_(dev, non_blocking=True)
            torch.cuda.synchronize()
            results["solo"] = {"d2h": 10 * n / (time.perf_counter() - t0) / 1e9}
            print(f"rank {rank} solo d2h {results['solo']['d2h']:.1f} GB/s", flush=True)
            del host
    mine = dict(rank=rank, gpu=local, bus=busid, numa_node=node, results=results)
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        conc = [r["results"]["default placement"] for r in allr]
        summary = dict(
            n_ranks=world, host_cpus=os.cpu_count(), numa_nodes=nodes, bytes_per_copy=n, copies=10,
            what="plain pinned cudaMemcpyAsync ceilings, all ranks copying at the same time (1 GiB buffers, 10 copies, wall "
                 "clock around a device synchronize) and one rank at a time (solo)",
            concurrent_d2h_gbs_per_rank=[c["d2h"] for c in conc], concurrent_h2d_gbs_per_rank=[c["h2d"] for c in conc],
            concurrent_d2h_duplex_gbs_per_rank=[c["d2h_duplex"] for c in conc],
            concurrent_d2h_gbs_total=sum(c["d2h"] for c in conc), concurrent_h2d_gbs_total=sum(c["h2d"] for c in conc),
            solo_d2h_gbs_per_rank=[r["results"]["solo"]["d2h"] for r in allr], ranks=allr)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(summary, open(f"gpurun_out/r2_numa_probe_n{world}.json", "w"), indent=1)
        print(json.dumps({k: v for k, v in summary.items() if k != "ranks"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
