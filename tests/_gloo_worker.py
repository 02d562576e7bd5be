"""Worker of tests/test_sharding_cpu.py: one rank of a world_size-2 gloo job (run as a subprocess)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from cwsl_digi_b200 import sharding


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = port
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.receivers_of_rank(5, rank, world)
    audio = torch.full((240000,), 100 * (mine[0] + 1), dtype=torch.int16)   # stand-in for one channel's slot audio
    audio[179968:] = 0
    got = sharding.gather_slot_audio(audio, dst=0)
    t = sharding.max_over_ranks(1.0 + rank)
    out = dict(rank=rank, mine=mine, tmax=t, first=[int(g[0]) for g in got] if got else None,
               last=[int(g[-1]) for g in got] if got else None)
    dist.barrier()
    dist.destroy_process_group()
    print("RESULT " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
