"""Host-side logic of bench.py (no GPU): the reference arm's CPU leg prints exactly one JSON line with the contract's keys,
the algorithmic-byte model matches SURVEY 8(d), and the product arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_algorithmic_bytes_model():
    import bench
    P, V, R, HW, M = 500_000, 399_983, 7_715_524, 1920 * 1080, 15
    A1, per = bench.algorithmic_bytes(P, V, R, HW, M)
    assert A1 == 44 * P + 776 * V + 244 * R + 84 * HW            # deg 3 @1080p, 6 sort passes
    assert per["render_bwd"] == 40 * R + 32 * HW + 36 * V and per["sort"] == 152 * R
    assert bench.physical_cores() >= 1


def test_reference_arm_cpu_leg_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "1500"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("Gaussians/sec") and d["unit"] == "Gaussians/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0


def test_product_arm_needs_a_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"], capture_output=True,
                         text=True, cwd=ROOT, timeout=600)
    assert out.returncode != 0, "bench.py must not produce a number without the CUDA path"
    assert not any(l.strip().startswith("{") for l in out.stdout.splitlines())
