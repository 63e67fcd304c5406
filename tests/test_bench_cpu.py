"""bench.py's reference arm (`--impl reference`) needs no GPU: the line the driver parses is checked here, on CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` times the reference's own CPU implementation of the path (oracle/_ref when the reference
    was copied here by build(), else the torch-CPU port) on a bounded sample and prints ONE JSON line with the same metric /
    config as the GPU arm, `impl: "reference"`, a `cpu_baseline` describing the run and an `e2e` that repeats the value with
    zero host<->device bytes."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    if "unavailable" in d:                       # allowed by the contract, never expected here (the oracle always exists)
        raise AssertionError(d)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["metric"].startswith("SAC gradient-updates/sec") and d["unit"] == "updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1.0
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
