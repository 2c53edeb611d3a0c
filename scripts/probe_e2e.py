"""Development probe: end-to-end host-array sweep (pdp_sweep_host) vs chunk count, pinned buffers, cfg2."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyro_b200 import problem
from pyro_b200.engine import Engine
from tests.cases import build_case

case = dict(system="SinglePendulum", x_grid_dim=[1001, 1001], u_grid_dim=[201], xbar=[-3.14, 0.0], INF=300.0)
_, grid, cf = build_case(case)
P = problem.extract(grid, cf, 1.0)
N = P.N
Jin = torch.empty(N, dtype=torch.float64).pin_memory(); Jo = torch.empty(N, dtype=torch.float64).pin_memory()
pio = torch.empty(N, dtype=torch.int64).pin_memory()
Jin.copy_(torch.from_numpy(np.random.default_rng(0).uniform(0, 250, N)))
mode = os.environ.get("PROBE_STREAM", "own")   # own | null | torch
if mode == "torch":
    torch.cuda.set_stream(torch.cuda.Stream())
for chunks in [int(c) for c in (sys.argv[1:] or [1, 2, 3, 4, 6, 8, 12, 16])]:
    os.environ["PYRODP_HOST_CHUNKS"] = str(chunks)
    eng = Engine(P)
    if mode != "own":
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        eng.sweep_host(Jin.numpy(), Jo.numpy(), pio.numpy())
    t0 = time.perf_counter(); K = 20
    for _ in range(K):
        eng.sweep_host(Jin.numpy(), Jo.numpy(), pio.numpy())
    dt = (time.perf_counter() - t0) / K
    print(json.dumps({"stream": mode, "chunks": chunks, "ms_per_call": round(dt * 1e3, 4), "evals_per_s": N * P.A / dt}), flush=True)
    eng.close()
