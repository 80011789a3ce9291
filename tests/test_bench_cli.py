"""bench.py on a box without a GPU: the reference arm runs (CPU oracle port, bounded sample) and prints the contract's
JSON line; the product arm refuses to run -- there is no CPU fallback."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600, env=env)


def test_reference_arm_prints_the_contract_line():
    res = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--sample", "2")
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "heatmaps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_product_arm_refuses_to_run_without_a_gpu():
    res = _run("--steps", "1", "--warmup", "1")
    assert res.returncode != 0
    assert "no CPU fallback" in res.stderr


def test_reference_arm_under_torchrun_prints_once():
    """Launched like the driver launches N > 1 (torch.distributed.run, one process per GPU): rank 0 alone runs the CPU
    arm and prints the line, the other ranks exit 0 without work."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
           "--sample", "2"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["value"] > 0
