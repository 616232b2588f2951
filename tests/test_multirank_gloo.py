"""world_size-2 gloo test (CPU) of the only collective on the path: the all-reduce(sum) of the
episode/lock metric vector, plus the env sharding arithmetic used with it."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]


def _worker(rank: int, world: int, port: int, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(REPO))
    from dl_reference_models_b200 import metrics

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = metrics.shard_range(10, rank, world)
        vec = torch.zeros(16, dtype=torch.float64)
        vec[0] = hi - lo                 # episodes: one per env of the shard
        vec[1] = float(sum(range(lo, hi)))  # return_sum: global env ids
        vec[3] = 1.0 if rank == 0 else 0.0
        out = metrics.allreduce_metrics(vec)
        q.put((rank, out["episodes"], out["return_sum"], out["return_mean"], out["success_rate"], (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_metric_allreduce_world_size_2():
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
    assert [r[5] for r in res] == [(0, 5), (5, 10)]
    for _rank, episodes, rsum, rmean, succ, _span in res:  # every rank holds the same global result
        assert episodes == 10 and rsum == 45.0 and rmean == 4.5 and succ == pytest.approx(0.1)


def test_single_process_allreduce_is_identity():
    import torch

    from dl_reference_models_b200 import metrics

    v = torch.arange(16, dtype=torch.float64)
    assert metrics.allreduce_metrics(v, 1)["return_sum"] == 1.0
