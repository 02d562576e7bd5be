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


def test_channel_slices_cover_every_channel_once():
    for n_rx, n_ch, w in [(1, 1024, 8), (1, 7, 8), (2, 100, 8), (3, 1000, 8), (8, 56, 8), (64, 1024, 8), (1, 33, 2),
                          (5, 64, 4), (1, 1, 1)]:
        seen = {}
        for r in range(w):
            for rx, lo, hi in sharding.channel_slices_of_rank(n_rx, n_ch, r, w):
                assert 0 <= lo < hi <= n_ch
                assert lo % 32 == 0
                for c in range(lo, hi):
                    assert (rx, c) not in seen
                    seen[(rx, c)] = r
        assert len(seen) == n_rx * n_ch
    # fewer receivers than ranks: every rank of a big receiver gets work, slices are balanced to a granule
    sizes = [sum(hi - lo for _, lo, hi in sharding.channel_slices_of_rank(1, 1024, r, 8)) for r in range(8)]
    assert sizes == [128] * 8
    assert sharding.channel_slices_of_rank(64, 1024, 3, 8) == [(r, 0, 1024) for r in range(3, 64, 8)]


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
