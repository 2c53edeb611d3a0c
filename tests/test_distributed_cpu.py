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


def _worker(rank, world, port, name, n_sweeps, mode, q, bounds=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pyro_b200 import distributed
        from tests.fake_engine import FakeEngine
        case = CASES[name]
        _, grid, cf = build_case(case)
        eng = distributed.ShardedEngine(grid, cf, case.get("alpha", 1.0), engine_factory=FakeEngine, mode=mode, bounds=bounds)
        eng.eval_terminal_cost()
        stats = eng.sweep(n_sweeps)
        J, pi, Jn = eng.get_J(), eng.get_pi(), eng.get_J_next()
        held = (eng.alloc_end - eng.alloc_begin, eng.halo_lo, eng.halo_hi)
        q.put((rank, J, pi, stats, (eng.begin, eng.end), eng.mode, held, Jn))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world,k,mode,expect,bounds", [
    ("pend_51x51x11", 2, 10, None, "halo", None),          # neighbour send/recv of halo planes
    ("dpend_example", 2, 2, None, "halo", None),
    ("cartpole_swingup", 3, 2, None, "halo", None),        # middle rank exchanges on both sides
    ("pend_51x51x11", 3, 2, "allgather", "allgather", None),  # forced fallback: in-place all-gather of whole slabs
    ("twolink_9", 8, 1, None, "halo", None),               # 9 planes on 8 ranks, one-plane halos: still neighbour exchange
    ("dpend_example", 6, 1, None, "allgather", None),      # halo (2 planes) wider than a slab (1): automatic fallback
    ("pend_51x51x11", 3, 2, None, "halo", [0, 9, 31, 51]),  # work-balanced (unequal) slabs, ShardedEngine.balanced's layout
])
def test_sharded_sweeps_equal_single_rank_goldens(name, world, k, mode, expect, bounds):
    gold = load_golden(name)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, k, mode, q, bounds)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    slabs = sorted(r[4] for r in results)
    assert slabs[0][0] == 0 and slabs[-1][1] == CASES[name]["x_grid_dim"][0]
    if bounds is not None:
        assert [s[0] for s in slabs] + [slabs[-1][1]] == bounds
    n0 = CASES[name]["x_grid_dim"][0]
    for rank, J, pi, stats, slab, got_mode, held, Jn in results:
        assert got_mode == expect
        if expect == "halo":   # memory per rank is slab + halo, not the whole grid
            assert held[0] <= (slab[1] - slab[0]) + held[1] + held[2] and (world == 1 or held[0] < n0)
        if f"J_{k-1}" in gold.files:
            assert np.array_equal(Jn, gold[f"J_{k-1}"])
        assert np.array_equal(J, gold[f"J_{k}"]), f"rank {rank}: J differs"
        assert np.array_equal(pi, gold[f"pi_{k}"]), f"rank {rank}: pi differs"
        d = gold[f"J_{k}"] - (gold[f"J_{k-1}"] if f"J_{k-1}" in gold.files else J * np.nan)
        assert stats.shape == (k, 3) and stats[-1, 0] == gold[f"J_{k}"].max()
        if not np.isnan(d).any():
            assert stats[-1, 1] == d.max() and stats[-1, 2] == d.min()


def test_rebalanced_bounds_properties():
    from pyro_b200.distributed import rebalanced_bounds
    equal = [r * 201 // 8 for r in range(8)] + [201]
    times = [2000, 2100, 2200, 2300, 2350, 2370, 2370, 2300]       # ms per rank with equal slabs (cfg5 on 8 GPUs, shape of r02f)
    new = rebalanced_bounds(equal, times, 21)
    assert new[0] == 0 and new[-1] == 201 and all(new[r + 1] - new[r] >= 21 for r in range(8))
    dens = np.concatenate([np.full(equal[r + 1] - equal[r], times[r] / (equal[r + 1] - equal[r])) for r in range(8)])
    work = [dens[new[r]:new[r + 1]].sum() for r in range(8)]
    assert max(work) < max(times) and max(work) / (sum(times) / 8) < 1.03      # within 3 % of perfectly even under the model
    assert rebalanced_bounds(equal, [1.0] * 8, 21) == equal                      # nothing to gain: unchanged
    assert rebalanced_bounds([0, 3, 6, 9], [1, 50, 1], 3) == [0, 3, 6, 9]        # minimum thickness leaves no room: unchanged
