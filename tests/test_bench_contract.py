"""CPU-side checks of bench.py's contract: the reference arm (the C oracle timed on the host
cores) runs without a GPU and prints one JSON line with the keys the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--size", "120", "64"], capture_output=True, text=True,
                       timeout=300, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"].startswith("Mcell-steps/s") and d["unit"] == "Mcell-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "c5-sor-dominated-channel"


def test_default_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm must fail, not fall back."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0",
                        "--size", "64", "32", "--no-cpu"], capture_output=True, text=True,
                       timeout=300, cwd=str(ROOT))
    assert r.returncode != 0
    assert "impl" not in r.stdout
