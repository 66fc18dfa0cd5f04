"""CPU: bench.py's reference arm (the reference's CPU path restated in torch fp64, timed on the host cores) produces the
contract's JSON line, with a `config` object identical to the one the B200 arm prints for the same flags."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *flags],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_json_contract():
    line = _run("--steps", "2", "--warmup", "0", "--size", "400")
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "steps_run", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "exact_gp_fit_predict_gflops" and line["unit"] == "GF/s"
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["steps"] == 2 and 1 <= line["steps_run"] <= 2
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert "torch fp64" in line["cpu_baseline"]["sample"] and "cholesky_ex" in line["cpu_baseline"]["sample"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"] > 0


def test_reference_arm_config_equals_the_b200_arm_config():
    """The driver compares the two arms' `config`: it must describe the workload only (nothing measured, no sample size)."""
    sys.path.insert(0, ROOT)
    import bench
    line = _run("--steps", "1", "--warmup", "0", "--size", "300", "--gpus", "1")
    args = argparse.Namespace(n=300, workload="per_gpu")
    assert line["config"] == bench.workload_config(args, 1)
    assert set(line["config"]) == {"workload", "n", "m_query", "kernel", "l2_policy"}
    # the default size is BASELINE configs[1] and the reference arm then really runs N = 40 000
    d = bench.workload_config(argparse.Namespace(n=40000, workload="per_gpu"), 1)
    assert "N=40000" in d["workload"] and "configs[1]" in d["workload"]
    assert "configs[3]" in bench.workload_config(argparse.Namespace(n=40000, workload="per_gpu"), 8)["workload"]
    assert "configs[4]" in bench.workload_config(argparse.Namespace(n=200000, workload="sharded"), 8)["workload"]
    assert "configs[2]" in bench.workload_config(argparse.Namespace(n=80000, workload="train"), 1)["workload"]


def test_flop_model():
    sys.path.insert(0, ROOT)
    import bench
    n, m = 40000, 300
    assert abs(bench.algorithmic_flops(n, m) - (n ** 3 / 3 + n * n * m + 2 * n * n)) < 1.0
    r = bench.step_roofline(200000, 4.0, 8, True)
    assert r["bound"] == "tensor" and 0.3 < r["frac"] < 1.0 and r["fp64_equivalent_over_dmma_peak"] > 1.0
    assert bench.OZ_PAIRS == 28


def test_roofline_inputs_come_from_committed_measurements(tmp_path):
    """`roofline.peak` and `roofline.traffic` are read from files measured on the pool (tools/microbench/i8_peak.cu, the ncu
    launch list), never hard-coded: the int8 peak is the N=256 tcgen05 microbenchmark, the traffic is dram bytes per launch
    of the dominant kernel as summarised by tools/summarize_launches.py --traffic-json."""
    sys.path.insert(0, ROOT)
    import bench
    pk = bench.int8_peak()
    assert "DERIVED" not in pk["source"] and 3000.0 < pk["sustained"] <= pk["burst"] < 5000.0
    t = bench.ncu_traffic("oz_mma_kernel")
    assert t is not None and t["launches"] > 50 and 1e8 < t["dram_bytes_per_launch"] < 1e10 and "ncu" in t["source"]
    # the summariser itself on a three-launch list
    csv = tmp_path / "l.csv"
    rows = ['"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"']
    for i, (name, ns, rd, wr) in enumerate([("void bgp::oz_mma_kernel<(int)1>(OzArgs)", "1000", "2000", "500"),
                                            ("void bgp::oz_mma_kernel<(int)1>(OzArgs)", "3000", "4000", "1500"),
                                            ("bgp::oz_slice_kernel(double)", "10", "64", "56")]):
        for metric, unit, val in (("dram__bytes_read.sum", "byte", rd), ("dram__bytes_write.sum", "byte", wr), ("gpu__time_duration.sum", "ns", ns)):
            rows.append(f'"{i}","1","python","h","{name}","1","7","(224, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","{metric}","{unit}","{val}"')
    csv.write_text("\n".join(rows) + "\n")
    out = tmp_path / "t.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"), str(csv), "--traffic-json", str(out)],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    d = json.loads(out.read_text())
    assert d["oz_mma_kernel"]["launches"] == 2 and d["oz_mma_kernel"]["dram_bytes_per_launch"] == 4000.0
    assert "oz_mma_kernel" in r.stdout and "launches=    2" in r.stdout
