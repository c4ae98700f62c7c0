"""bench.py contract checks that need no GPU: the reference arm falls back to the CPU oracle port when no CUDA device is
present (this container), prints ONE JSON line with the keys the driver reads, and non-zero ranks of a torchrun launch exit 0
without work."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


@pytest.mark.skipif(torch.cuda.is_available(), reason="on a GPU box the reference arm runs the reference CUDA kernels (covered by the round-end bench)")
def test_reference_arm_line_on_cpu():
    out = _run(None, "--impl", "reference", "--steps", "1", "--warmup", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "voxels/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "sample" in d["cpu_baseline"] and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    out = _run({"WORLD_SIZE": "2", "RANK": "1", "LOCAL_RANK": "1", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29998"}, "--impl", "reference", "--gpus", "2",
               "--steps", "1", "--warmup", "1")
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_fails_loudly_without_a_gpu():
    out = _run(None, "--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert out.returncode != 0 and "CUDA" in (out.stderr + out.stdout)


def test_clock_sampler_reports_unavailable_without_a_gpu():
    sys.path.insert(0, ROOT)
    import bench
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    c = bench.ClockSampler(0)
    c.start()
    r = c.stop()
    assert r["sm_mhz"] is None and r["reasons"]
