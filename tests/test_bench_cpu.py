"""CPU: bench.py's reference arm (the CPU restatement timed on the host cores) produces the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-n", "400"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "exact_gp_fit_predict_gflops" and line["unit"] == "GF/s"
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"] > 0


def test_flop_model():
    sys.path.insert(0, ROOT)
    import bench
    n, m = 40000, 300
    assert abs(bench.algorithmic_flops(n, m) - (n ** 3 / 3 + n * n * m + 2 * n * n)) < 1.0
    r = bench.step_roofline(200000, 5.79, 8, True)
    assert r["bound"] == "tensor" and 0.3 < r["frac"] < 1.0 and r["fp64_equivalent_over_dmma_peak"] > 1.0
