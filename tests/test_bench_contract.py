"""bench.py on the CPU: the reference arm (`--impl reference`: the unmodified reference's CPUSolver through
oracle/_ref/ref_driver, or the oracle port when that binary is absent) prints the contract's JSON line from rank 0 and
nothing from the other ranks; the GPU arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def run(args, **env):
    e = dict(os.environ, **{k: str(v) for k, v in env.items()})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    out = run(["--impl", "reference", "--workload", "pin-cell", "--steps", "2", "--warmup", "3", "--gpus", "1"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "segment-group integrations/s" and d["unit"] == "integrations/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"] == "pin-cell" and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0                               # nothing of this repo's kernels runs in this arm


def test_reference_arm_is_silent_on_other_ranks():
    out = run(["--impl", "reference", "--workload", "pin-cell", "--steps", "1", "--gpus", "2"], RANK=1, WORLD_SIZE=2,
              LOCAL_RANK=1)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = run(["--steps", "1", "--workload", "pin-cell", "--also", "none"])
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
