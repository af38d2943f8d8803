"""The GPU arm of bench.py prints one line with the keys the driver and the judge read."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3",
                          "--no-cpu-baseline"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert k in d, k
    assert d["steps"] == 3 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["gpu_launches"] >= 3 * 5      # theta->Trels, fused forward, fused adjoint, R->G, G.B
    r = d["roofline"]
    assert r["bound"] == "fp32" and 0 < r["frac"] < 1.2 and r["unit"] == "TFLOP/s" and 50 < r["peak"] < 90
    assert d["roofline_interp"]["bound"] == "hbm" and 0 < d["roofline_interp"]["frac"] < 1.2
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 64 * 34 * 4 + 64 * 256 * 256 * 4 and e["d2h_bytes_per_step"] > 0
