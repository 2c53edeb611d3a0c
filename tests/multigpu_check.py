"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py

Every rank drives its slab through ``pyro_b200.distributed.ShardedEngine`` (NCCL) and the gathered
J / pi must equal the reference goldens bit for bit, in every exchange mode: halo send/recv with
boundary-first overlap, halo send/recv without overlap, and the all-gather fallback.
tests/test_multigpu.py launches this when the box has >= 2 GPUs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from pyro_b200 import distributed, problem
    from pyro_b200.engine import Engine
    from tests.cases import CASES, build_case
    from tests.conftest import load_golden

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    failures = 0
    for name in ("pend_51x51x11", "pend_101x101x21", "pend_time_41x61x7", "dpend_example", "twolink_soft", "cartpole_swingup"):
        case, gold = CASES[name], load_golden(name)
        _, grid, cf = build_case(case)
        k = case["snapshots"][1]
        for mode, overlap, backend, halo in (("halo", True, "native", "peer"), ("halo", False, "native", "peer"), ("halo", True, "native", "nccl"),
                                             ("halo", False, "native", "nccl"), ("allgather", False, "native", None),
                                             ("halo", True, "torch", None), ("allgather", False, "torch", None)):
            try:
                eng = distributed.ShardedEngine(grid, cf, case.get("alpha", 1.0), mode=mode, overlap=overlap, backend=backend, halo=halo)
            except ValueError:
                if rank == 0:
                    print(f"[multigpu] {name} {mode}: halo does not fit {world} ranks, skipped")
                continue
            eng.eval_terminal_cost()
            stats = eng.sweep(k)
            J, pi, Jn = eng.get_J(), eng.get_pi(), eng.get_J_next()
            ok = np.array_equal(J, gold[f"J_{k}"]) and np.array_equal(pi, gold[f"pi_{k}"])
            d = J - Jn
            ok = ok and stats[-1, 0] == J.max() and stats[-1, 1] == d.max() and stats[-1, 2] == d.min()
            held = eng.alloc_end - eng.alloc_begin
            if rank == 0:
                print(f"[multigpu] {name} W={world} {eng.backend} mode={eng.mode} halo={getattr(eng, 'halo', '-')} overlap={eng.overlap} k={k} planes held {held}/{eng.n0} "
                      f"halo=({eng.halo_lo},{eng.halo_hi}): {'OK' if ok else 'MISMATCH'}", flush=True)
            failures += 0 if ok else 1
            eng.close()
    # mid-size random-J check against a single-GPU engine on the same device (rough J, every corner weight matters)
    case = dict(system="CartPole", x_grid_dim=[33, 21, 19, 23], u_grid_dim=[9], xbar=[0.0, float(np.pi), 0.0, 0.0], INF=1000.0)
    _, grid, cf = build_case(case)
    J0 = np.random.default_rng(7).uniform(0, 300, grid.nodes_n)
    single = Engine(problem.extract(grid, cf, 1.0))
    single.set_J(J0)
    single.sweep(3)
    Jref, piref = single.get_J(), single.get_pi()
    single.close()
    for overlap, halo in ((True, "peer"), (False, "peer"), (True, "nccl"), (False, "nccl")):
        eng = distributed.ShardedEngine(grid, cf, 1.0, overlap=overlap, halo=halo)
        eng.set_J(J0)
        eng.sweep_nowait()            # the non-blocking form the benchmark uses
        eng.sweep_nowait()
        assert eng.collect_stats().shape == (2, 3)
        eng.sweep(1)
        ok = np.array_equal(eng.get_J(), Jref) and np.array_equal(eng.get_pi(), piref)
        if rank == 0:
            print(f"[multigpu] cartpole 33x21x19x23 random J W={world} mode={eng.mode} halo={eng.halo} overlap={eng.overlap}: {'OK' if ok else 'MISMATCH'}", flush=True)
        failures += 0 if ok else 1
        # clean_infeasible_set rewrites slab nodes; the neighbours' halo copies must follow (exchange of the CURRENT J)
        eng.clean_infeasible_set(1.0, 3)
        eng.sweep(1)
        Jc = eng.get_J()
        ref = Engine(problem.extract(grid, cf, 1.0))
        ref.set_J(J0); ref.sweep(3); ref.clean_infeasible_set(1.0, 3); ref.sweep(1)
        ok = np.array_equal(Jc, ref.get_J())
        ref.close()
        if rank == 0:
            print(f"[multigpu] ... clean_infeasible_set + sweep, halo={eng.halo}: {'OK' if ok else 'MISMATCH'}", flush=True)
        failures += 0 if ok else 1
        eng.close()
    # peer-store exchange with an ASYMMETRIC halo (pend_time: 4 rows below, 5 above) and the interior kernel held back by
    # a spin kernel: a neighbour racing ahead into this rank's halo planes would corrupt interior reads (VERDICT r01 weak #3;
    # fixed by boundaries of max(halo_lo, halo_hi) planes in peer mode)
    os.environ["PYRODP_TEST_INTERIOR_DELAY_US"] = "2000"
    for name in ("pend_time_41x61x7", "pend_101x101x21"):
        case, gold = CASES[name], load_golden(name)
        _, grid, cf = build_case(case)
        k = case["snapshots"][2]
        try:
            eng = distributed.ShardedEngine(grid, cf, case.get("alpha", 1.0), mode="halo", overlap=True, backend="native", halo="peer")
        except ValueError:
            continue
        eng.eval_terminal_cost()
        for _ in range(k):
            eng.sweep_nowait()
        eng.collect_stats()
        ok = np.array_equal(eng.get_J(), gold[f"J_{k}"]) and np.array_equal(eng.get_pi(), gold[f"pi_{k}"])
        if rank == 0:
            print(f"[multigpu] {name} peer stores, delayed interior, halo=({eng.halo_lo},{eng.halo_hi}), {k} sweeps back to back: "
                  f"{'OK' if ok else 'MISMATCH'}", flush=True)
        failures += 0 if ok else 1
        eng.close()
    os.environ.pop("PYRODP_TEST_INTERIOR_DELAY_US")

    # BASELINE config 4 AT FULL SIZE (CartPole 151^4 x 51): every rank's slab after 3 sharded sweeps must hash to the same
    # bytes as the corresponding planes of a one-GPU run
    if os.environ.get("MULTIGPU_FULL", "1") != "0":
        import hashlib
        from bench import WORKLOADS
        _, grid, cf = build_case(WORKLOADS["cfg4"])
        eng = distributed.ShardedEngine(grid, cf, 1.0)
        eng.eval_terminal_cost()
        eng.sweep(3)
        keng = eng.eng
        lo, cnt = keng.slab_begin * keng.plane, keng.slab_nodes
        mine = (hashlib.sha256(keng.get_range("J", lo, cnt).tobytes()).hexdigest(), hashlib.sha256(keng.get_range("pi", lo, cnt).tobytes()).hexdigest(), lo, cnt)
        eng.close()
        torch.cuda.empty_cache()
        every = [None] * world
        dist.all_gather_object(every, mine)
        if rank == 0:
            single = Engine(problem.extract(grid, cf, 1.0))
            single.eval_terminal_cost()
            single.sweep(3)
            ok = True
            for r, (hj, hp, lo_r, cnt_r) in enumerate(every):
                ok = ok and hj == hashlib.sha256(single.get_range("J", lo_r, cnt_r).tobytes()).hexdigest()
                ok = ok and hp == hashlib.sha256(single.get_range("pi", lo_r, cnt_r).tobytes()).hexdigest()
            single.close()
            print(f"[multigpu] cfg4 CartPole 151^4 x 51 FULL SIZE, {world} slabs, 3 sweeps: sha256 of every slab's J and pi vs one GPU: "
                  f"{'OK' if ok else 'MISMATCH'}", flush=True)
            failures += 0 if ok else 1

    t = torch.tensor([failures], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print(f"[multigpu] {'ALL OK' if t.item() == 0 else 'FAILURES: %d' % t.item()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
