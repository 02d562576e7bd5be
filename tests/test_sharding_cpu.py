"""world_size-2 gloo test of the multi-GPU host logic: receiver partition, station-level gather of
per-slot audio, max-over-ranks timing. (The data path itself has no collective.)"""
import json
import os
import socket
import subprocess
import sys

import pytest

from cwsl_digi_b200 import sharding


def test_partition_covers_every_receiver_once():
    for n, w in [(64, 1), (64, 2), (64, 4), (64, 8), (8, 8), (7, 4), (3, 8)]:
        seen = []
        for r in range(w):
            mine = sharding.receivers_of_rank(n, r, w)
            assert all(sharding.owner_of_receiver(x, w) == r for x in mine)
            seen += mine
        assert sorted(seen) == list(range(n))
        sizes = [len(sharding.receivers_of_rank(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.receivers_of_rank(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gather_gloo():
    world, port = 2, _free_port()
    worker = os.path.join(os.path.dirname(__file__), "_gloo_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), str(world), str(port)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    res = []
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out[-2000:]
        line = [l for l in out.splitlines() if l.startswith("RESULT ")][-1]
        res.append(json.loads(line[7:]))
    root = [r for r in res if r["first"] is not None][0]
    assert root["rank"] == 0 and root["first"] == [100, 200] and root["last"] == [0, 0]
    assert all(abs(r["tmax"] - 2.0) < 1e-12 for r in res)          # max over ranks
    assert sorted(sum((r["mine"] for r in res), [])) == [0, 1, 2, 3, 4]
