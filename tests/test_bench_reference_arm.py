"""The CPU arm of bench.py (`--impl reference`) needs no GPU: it must print one JSON line with the contract's keys for
the default workload (bounded sample) and for the whole-model mode of configs[0]."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_reference_arm_default_workload_small_sample():
    # a reduced shape keeps the test short; the keys and the arithmetic are those of the real run
    d = _run("--height", "128", "--width", "128", "--batch", "2", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["gpu_launches"] == 0


def test_reference_arm_whole_model_mode():
    d = _run("--config", "0", "--mode", "infer", "--height", "128", "--width", "128", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and "full model" in d["metric"]
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"


def test_reference_arm_has_no_collective_to_time():
    d = _run("--mode", "allreduce")
    assert d["impl"] == "reference" and "unavailable" in d
