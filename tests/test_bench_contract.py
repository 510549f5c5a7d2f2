"""bench.py's command-line contract, as far as it can be checked without a GPU: the reference arm (the oracle timed on the host
cores) prints ONE JSON line with the keys the driver reads, and our own arm fails loudly on a machine without a CUDA device — the
product path has no CPU fallback."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--size", "10", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "SIMPLE iters/s" and d["unit"] == "iter/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] / 1e3 - 1.0) < 1e-9 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and cb["unit"] == "iter/s" and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    assert set(d["phases_ms_per_step"]) == {"momentum_assembly", "momentum_solves", "pressure_assembly", "pressure_solve", "correction"}


def test_reference_arm_runs_on_rank_zero_only():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--size", "10", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.returncode, r.stdout[-500:], r.stderr[-500:])


def test_our_arm_fails_loudly_without_a_gpu():
    import orc_b200
    try:
        orc_b200.default_context(0)
    except orc_b200.OrcError:
        pass
    else:
        pytest.skip("a CUDA device is present")
    r = _run("--size", "8", "--steps", "1", "--warmup", "0", "--no-e2e", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr or "NVIDIA" in r.stderr            # the library's own refusal, or torch's on the way to it
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines())     # no result line of any kind
