"""CPU: the committed bench lines (profiles/rNN_bench_*.json, produced by bench.py on a B200) carry every key of the
bench contract, and bench.py's static pieces (argument parser, metric / unit) match BASELINE.json."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"}


def load(name):
    """The newest committed round's file of that name (profiles/r02_... before profiles/r01_...)."""
    for rnd in ("r02", "r01"):
        path = os.path.join(ROOT, "profiles", f"{rnd}_{name}")
        if os.path.exists(path):
            return json.load(open(path))
    pytest.skip(f"{name} not committed yet")


def test_one_gpu_line_has_the_contract_keys():
    d = load("bench_1gpu.json")
    assert BASE_KEYS <= set(d)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["metric"].split()[0] in base["metric"] or "frames/s" in base["metric"]
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor", "fp32")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference") and c["value"] > 0
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["gpu_launches"] == len(d["kernel_ms"]) * d["steps"]
    k = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(k)
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(k["reasons"]))
    assert len(d["kernel_ms"]) in (15, 16) and d["dominant_kernel"] in d["kernel_ms"]


def test_reference_arm_line():
    d = load("bench_reference_arm.json")
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_eight_gpu_line_scales():
    one, eight = load("bench_1gpu.json"), load("bench_8gpu.json")
    assert eight["n_gpus"] == 8 and eight["config"]["global_batch"] == 8 * one["config"]["global_batch"]
    assert eight["value"] > 5 * one["value"]


def test_bench_cli_help_lists_the_contract_flags():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout


def test_reference_arm_under_torchrun_prints_one_line_from_rank_zero():
    """N > 1: the driver launches the reference arm like the GPU arm; rank 0 alone runs and prints, the other rank exits 0."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
