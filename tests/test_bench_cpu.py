"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`: the
reference's own CPU implementation of the path on all host threads, a bounded sample of the workload) prints ONE
JSON line with the keys the driver reads, and the product arm refuses to run without a device instead of falling
back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-sample-mib", "8", "--strings", "300"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["dtype"] == "u8" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--size-gib", "0.01",
                        "--strings", "300", "--no-e2e", "--no-cpu", "--no-parity"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode != 0                      # no CPU fallback: the product arm fails loudly
    assert not [l for l in p.stdout.splitlines() if l.startswith("{") and '"value"' in l]
