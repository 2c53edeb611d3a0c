"""world_size-2 (and 3) gloo test of the slab/exchange host logic with CPU stand-in engines."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.cases import CASES, build_case
from tests.conftest import load_golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, n_sweeps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pyro_b200 import distributed
        from tests.fake_engine import FakeEngine
        case = CASES[name]
        _, grid, cf = build_case(case)
        eng = distributed.ShardedEngine(grid, cf, case.get("alpha", 1.0), engine_factory=FakeEngine)
        eng.eval_terminal_cost()
        stats = eng.sweep(n_sweeps)
        J, pi = eng.get_J(), eng.get_pi()
        q.put((rank, J, pi, stats, (eng.begin, eng.end)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world,k", [("pend_51x51x11", 2, 10), ("dpend_example", 2, 2), ("cartpole_swingup", 3, 2)])
def test_sharded_sweeps_equal_single_rank_goldens(name, world, k):
    gold = load_golden(name)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    slabs = sorted(r[4] for r in results)
    assert slabs[0][0] == 0 and slabs[-1][1] == CASES[name]["x_grid_dim"][0]
    for rank, J, pi, stats, _ in results:
        assert np.array_equal(J, gold[f"J_{k}"]), f"rank {rank}: J differs"
        assert np.array_equal(pi, gold[f"pi_{k}"]), f"rank {rank}: pi differs"
        d = gold[f"J_{k}"] - (gold[f"J_{k-1}"] if f"J_{k-1}" in gold.files else J * np.nan)
        assert stats.shape == (k, 3) and stats[-1, 0] == gold[f"J_{k}"].max()
        if not np.isnan(d).any():
            assert stats[-1, 1] == d.max() and stats[-1, 2] == d.min()
