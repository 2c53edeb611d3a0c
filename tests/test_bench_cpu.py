"""bench.py contract pieces that run without a GPU: the reference arm's JSON line, the helper-process clock sampler's
fallback, and the algorithmic-bytes figure of SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_per_eval_is_the_contract_figure():
    assert bench.b_eval(2, 201) == pytest.approx(32.0 + 24.0 / 201)
    assert bench.b_eval(4, 51) == pytest.approx(128.0 + 24.0 / 51)


def test_clock_sampler_without_nvml_reports_nulls_and_never_fails():
    s = bench.ClockSampler(0)
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_arm_prints_one_contract_line():
    """--impl reference: the reference's CPU path on the host cores (the unmodified pyro classes when /root/reference or
    baseline/_ref is there, else the oracle port), one JSON line."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "cfg1", "--no-extra"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "state_action_evals_per_s" and d["unit"] == "evals/s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["higher_is_better"] is True and d["gpu_launches"] == 0


def test_own_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a CUDA device the own arm exits with an error instead of timing something else."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except Exception:
        pytest.skip("torch unavailable")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_after_the_sweep_records_on_oracle_backed_engines(monkeypatch):
    """bench.py's rollout / spline records (N = 1) with the Engine replaced by a CPU stand-in — the C oracle for the fused
    sweeps and the table builder's values, the emulated kernels for the rollouts and the spline backups — at reduced sizes:
    checks the record's host logic and its in-run parity numbers without a GPU."""
    import numpy as np
    import torch
    import bench
    from pyro_b200 import dynamicprogramming
    from tests.emu import emu
    from tests.fake_engine import FakeEngine

    class StandIn(FakeEngine):
        kernel_info = "cpu stand-in"
        last_sweep_ms = 1.0

        def __init__(self, P):
            super().__init__(P)
            self.lut = self.interpolant = None

        def sweep(self, n_sweeps=1):
            if self.problem.system_id != 0:
                return super().sweep(n_sweeps)
            out = np.empty((n_sweeps, 3))
            for k in range(n_sweeps):
                Jn = self.get_J()
                assert self.interpolant == "spline3"
                J, pi, out[k], _ = emu.spline_sweep(self.problem, Jn, *self.lut)
                self._slab(1 - self.cur)[:] = torch.from_numpy(J)
                self.pi[:self.slab_nodes] = torch.from_numpy(pi)
                self.cur = 1 - self.cur
            return out

        def set_lut(self, x_next, G):
            self.lut = (np.array(x_next), np.array(G))

        def set_interpolant(self, which):
            self.interpolant = which

        def build_tables(self, node_begin=0, count=None, x_next=True, x_ok=True, G=True):
            xn, Gt = dynamicprogramming.build_lookup_tables(self.problem._grid, self.problem._cf, exact_inf=False)
            return xn, None, Gt

        def rollout(self, phys, x0, npts, dt, stride=1, with_inputs=True):
            return emu.rollout(self.problem, self.get_pi(), phys, x0, npts, dt, stride)

    made = []

    def factory(P):
        eng = StandIn(P)
        made.append(eng)
        return eng
    monkeypatch.setattr(dynamicprogramming, "Engine", factory)
    # the table builder of the stand-in needs the grid objects of the planner that asked: remember them at extraction
    real_extract = dynamicprogramming._problem.extract

    def extract(grid_sys, cf, *a, **k):
        P = real_extract(grid_sys, cf, *a, **k)
        P._grid, P._cf = grid_sys, cf
        return P
    monkeypatch.setattr(dynamicprogramming._problem, "extract", extract)
    small = dict(system="SinglePendulum", x_grid_dim=[21, 25], u_grid_dim=[5], xbar=[-3.14, 0.0], INF=300.0)
    rec = bench.after_the_sweep_records(n_traj=64, policy_dims=(31, 31), policy_sweeps=3, spline_case=small)
    roll, spl = rec["rollout_batches"], rec["spline_class"]
    assert "error" not in roll and "error" not in spl, rec
    assert roll["parity"]["policy_equal"] and roll["parity"]["x_Linf_error"] <= 1e-9 and roll["parity"]["u_Linf_error"] <= 1e-9
    assert roll["value"] > 0 and "64 trajectories" in roll["workload"]
    assert spl["parity"]["J_Linf_error"] <= 1e-9 and spl["parity"]["pi_mismatches"] == 0 and spl["value"] > 0
