"""bench.py's reference arm runs without a GPU (it times the oracle port on the host): check the JSON line it prints
against the bench contract (keys the driver reads), and that the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "seams_per_sec_4k_rgba" and line["unit"] == "seams/s"
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["higher_is_better"] is True
    assert line["value"] > 0 and abs(line["value"] - 200 / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"] + 1e-9
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]
    assert "3840x2160" in line["config"]["workload"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # the GPU box runs the real thing (bench.py itself); nothing to check here
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300, cwd=REPO)
    assert out.returncode != 0, "the product arm must fail loudly without a CUDA device (no CPU fallback)"
