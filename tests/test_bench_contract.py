"""bench.py's output contract where it can be checked without a GPU: the reference arm (CPU oracle port of the path) prints one
JSON line with the keys the driver reads, on the same metric / unit / config as the CUDA arm; the CUDA arm refuses to run
without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          env={**os.environ, **(env or {})}, cwd=ROOT)


def test_reference_arm_line_has_the_contract_keys():
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--edges", "200000")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "share_gather_edges_per_sec" and d["unit"] == "edges/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "u64"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_under_torchrun_env_only_rank0_prints():
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--edges", "100000",
                  env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_cuda_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        return
    r = run_bench("--steps", "1", "--warmup", "1", "--edges", "100000")
    assert r.returncode != 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert "no CPU fallback" in r.stderr or "CUDA" in r.stderr
