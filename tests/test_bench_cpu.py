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
