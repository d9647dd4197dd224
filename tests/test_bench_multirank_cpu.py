"""world_size-2 gloo test of bench.py's multi-rank aggregation (replicas only: no data-path
collective; the only exchange is barrier + max of the per-rank times)."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import bench
    w, r, local, dist = bench.dist_setup(world, backend="gloo")
    assert (w, r) == (world, rank)
    # rank r needs (r+1) seconds for 10 steps -> the job takes 2 s -> 2 ranks * 10 steps / 2 s
    slowest = bench.barrier_max(dist, local, float(rank + 1))
    rate = bench.aggregate_rate(dist, local, w, 10, float(rank + 1))
    q.put((rank, slowest, rate))
    dist.destroy_process_group()


def test_two_rank_gloo_aggregation():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, slowest, rate in res:
        assert slowest == 2.0
        assert rate == 10.0
