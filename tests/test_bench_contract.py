"""bench.py's output contract, checked on the CPU with the reference arm (`--impl reference` times the C/OpenMP port of the reference
loop, oracle/, on a bounded row sample: no GPU involved): exactly ONE line on stdout, a JSON object with the keys the driver reads."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    run = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    lines = [l for l in run.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, run.stdout[:500]  # library chatter (NCCL banners, ...) goes to stderr
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["ms_per_step_is_extrapolated"] is True
    assert "workload" in d["config"] and "MaternP(2)" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "rows" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_product_arm_fails_loudly_without_a_gpu():
    """the product has no CPU path: on a machine without a CUDA device bench.py must not print a result line"""
    import covfn_b200 as cf

    if cf.device_count() > 0:
        import pytest

        pytest.skip("a CUDA device is present")
    run = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "c1", "--steps", "1", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode != 0
    assert not [l for l in run.stdout.split("\n") if l.strip().startswith("{")]
