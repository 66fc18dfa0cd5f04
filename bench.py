#!/usr/bin/env python
"""bench.py -- exact-GP fit+predict throughput on B200 (BASELINE.json metric), see DESIGN.md "Measurement".

A *step* is one full pass of the hot path over one synthetic 8s1p data set (SURVEY.md 8d generator):
    build K (Wiener + RBF-ARD + noise)  ->  Cholesky  ->  alpha  ->  LML  ->  predict mean/var at M=300 queries.
N=1 GPU   : BASELINE.json configs[1]  (full_gp, N=40 000, D=3(+t), 1xB200).
N>1 GPUs  : configs[3]  (one independent N=40 000 GP per GPU, no collective; weak scaling) -- the reference's own
            process-per-GPU layout (gp_runner.py:246-298).  `--workload sharded` runs configs[4] instead (one GP
            block-row-sharded over all ranks with NCCL panel broadcasts).

value  = algorithmic GFLOP/s (N^3/3 + N^2 M + 2 N^2 per GP) with X, y resident in HBM.
e2e    = the same metric through the public API with HOST (pinned) inputs and host outputs inside the timed region.
--impl reference times the CPU restatement of the reference's GPyTorch Cholesky path (oracle/gp_oracle.py; GPyTorch
itself is not installable here, DESIGN.md "Reference arm") on the box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_QUERY = 300
NOISE = 2.33e-6
FP64_DMMA_PEAK_TFLOPS = 37.1   # measured on this pool's B200: profiles/fp64_peak_r01.txt (MEASURED_PEAKS.json has no fp64 entry)
# dram__bytes_read.sum + dram__bytes_write.sum summed over the kernels of one bgp_potrf call, from the committed ncu launch
# lists (profiles/launches_r01_final_summary.txt: int8 path 280 GB; launches_r01_summary.txt: DMMA-only path 315 GB) -- N -> bytes
TRAFFIC_BYTES_PER_POTRF = {40000: 2.8e11}


def algorithmic_flops(n: int, m: int = M_QUERY) -> float:
    return n ** 3 / 3.0 + float(n) ** 2 * m + 2.0 * float(n) ** 2


def step_roofline(n: int, sec: float, world: int, ozaki: bool) -> dict:
    """Roofline of a whole sharded step (N^3/3 of N^3/3 + N^2 M + 2 N^2 flop is the factorisation), per GPU."""
    fp64_equiv = (n ** 3 / 3.0) / sec * 1e-12 / world
    if not ozaki:
        return {"bound": "tensor", "achieved": fp64_equiv, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s per GPU (whole step)",
                "frac": fp64_equiv / FP64_DMMA_PEAK_TFLOPS, "traffic": None}
    try:
        bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        src = "2 x MEASURED_PEAKS.json bf16_tflops"
    except Exception:
        bf16, src = 1590.0, "2 x fallback 1.59 PFLOP/s"
    ach = 36.0 * fp64_equiv
    return {"bound": "tensor", "achieved": ach, "peak": 2.0 * bf16, "unit": "TOP/s int8 per GPU (whole step; 36 int8 ops per fp64 flop)",
            "frac": ach / (2.0 * bf16), "peak_source": src, "fp64_equivalent_tflops_per_gpu": fp64_equiv,
            "fp64_equivalent_over_dmma_peak": fp64_equiv / FP64_DMMA_PEAK_TFLOPS, "traffic": None}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------- reference arm
def run_reference(args, n_gpus: int):
    """CPU arm: oracle port of the reference's GPyTorch Cholesky path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import gp_oracle as orc
    threads = _all_host_threads()
    ns = args.ref_n
    x, y = orc.synth_field_data(ns, seed=0)
    xq = orc.query_grid(x)
    spec = orc.battgp_spec()

    def step():
        f = orc.fit(spec, x, y, NOISE)
        return orc.predict(spec, x, f, xq)

    for _ in range(args.warmup if args.warmup < 2 else 1):   # CPU warm-up is about page faults only
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    gf = algorithmic_flops(ns) / dt * 1e-9
    sample = f"fit+predict at N={ns} (same generator/hyper-parameters as the N={args.n} workload), fp64 numpy/LAPACK"
    line = {"impl": "reference", "metric": "exact_gp_fit_predict_gflops", "value": gf, "unit": "GF/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, n_gpus), "n": args.n, "m_query": M_QUERY, "kernel": "wiener+rbf_ard",
                       "sample_n": ns},
            "cpu_baseline": {"value": gf, "unit": "GF/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": gf, "unit": "GF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 for multi-rank launches; the CPU arm must still get every host core."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(n)
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return n


def workload_name(args, n_gpus):
    if args.workload == "sharded":
        return f"full_gp Wiener+RBF-ARD N={args.n} block-row-sharded Cholesky over {n_gpus} GPU(s) (BASELINE configs[4])"
    if n_gpus == 1:
        return f"full_gp Wiener+RBF-ARD N={args.n} D=3(+t) fit+predict on 1xB200 (BASELINE configs[1])"
    return f"8s1p per-cell batch: {n_gpus} independent full_gp N={args.n}, one per GPU, no collective (BASELINE configs[3])"


# ---------------------------------------------------------------------------------------------------- GPU arm
def cpu_baseline(args):
    import numpy as np  # noqa: F401
    from oracle import gp_oracle as orc
    threads = _all_host_threads()
    ns = args.ref_n
    x, y = orc.synth_field_data(ns, seed=0)
    xq = orc.query_grid(x)
    spec = orc.battgp_spec()
    t0 = time.perf_counter()
    f = orc.fit(spec, x, y, NOISE)
    orc.predict(spec, x, f, xq)
    dt = time.perf_counter() - t0
    return {"value": algorithmic_flops(ns) / dt * 1e-9, "unit": "GF/s", "cores": threads, "kind": "port",
            "sample": f"one fit+predict at N={ns} (same generator/hyper-parameters), {dt:.1f} s of fp64 numpy/LAPACK"}


def run_gpu(args, n_gpus: int):
    import numpy as np
    import torch
    import torch.distributed as dist
    from battgp_b200 import engine as E
    from battgp_b200.synth import query_grid, synth_field_data

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != n_gpus:
        raise SystemExit(f"--gpus {n_gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {n_gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; battgp_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.workload == "sharded":
        from battgp_b200 import sharded
        return sharded.bench(args, rank, world, dev)

    eng = E.get_engine(dev)
    n = args.n
    spec = E.battgp_spec()
    x_np, y_np = synth_field_data(n, seed=0, cell=rank if world > 1 else 0)
    xq_np = query_grid(x_np, M_QUERY)
    # device-resident inputs for `value`; pinned host copies for `e2e`
    xd, yd, xqd = (torch.tensor(a, device=dev) for a in (x_np, y_np, xq_np))
    xh, yh, xqh = (torch.tensor(a).pin_memory() for a in (x_np, y_np, xq_np))
    K = E.alloc_matrix(n + M_QUERY, n, dev)       # K with the M query rows appended (fused predictive solve)
    ev_p0, ev_p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    potrf_ms = []

    def step_device():
        st = E.fit(spec, xd, yd, NOISE, K_out=K, potrf_events=(ev_p0, ev_p1), xq=xqd)
        mean, var = E.predict(st, xqd)
        return st, mean, var

    def step_e2e():
        x = xh.to(dev, non_blocking=True); y = yh.to(dev, non_blocking=True); xq = xqh.to(dev, non_blocking=True)
        st = E.fit(spec, x, y, NOISE, K_out=K, xq=xq)
        mean, var = E.predict(st, xq)
        return mean.cpu(), var.cpu()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps, collect_potrf=False):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
            if collect_potrf:
                torch.cuda.synchronize(dev)       # fit() already synchronised inside bgp_potrf; this is free
                potrf_ms.append(ev_p0.elapsed_time(ev_p1))
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    ms_total = timed(step_device, args.steps, collect_potrf=True)
    launches = eng.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    # e2e: host buffers in, host results out
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # outside the timed region: parity evidence at the full size -- matrix-free residual and the oracle-free identity
    st_chk, mean_chk, var_chk = step_device()
    resid = E.residual(st_chk, yd)
    finite = bool(torch.isfinite(mean_chk).all() and torch.isfinite(var_chk).all() and (var_chk > 0).all())
    lml_chk = st_chk.lml
    del st_chk
    flops = algorithmic_flops(n)
    sec = ms_total / 1e3 / args.steps
    sec_e2e = ms_e2e / 1e3 / args.steps
    value = world * flops / sec * 1e-9
    e2e_value = world * flops / sec_e2e * 1e-9
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pm = statistics.mean(potrf_ms)
    fp64_equiv = n ** 3 / 3.0 / (pm * 1e-3) * 1e-12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
    bf16_src = "MEASURED_PEAKS.json bf16_tflops (burst)" if "bf16_tflops" in peaks else "fallback 1.59 PFLOP/s (B200_PROFILING.md)"
    if eng.ozaki:
        # the trailing updates run as 36 exact int8 x int8 -> int32 tcgen05 products per fp64 product (csrc/ozaki.cu);
        # the int8 tensor rate on sm_100a is 2x the bf16 rate, so the denominator is 2 x the measured bf16 peak
        int8_ops = 36.0 * n ** 3 / 3.0
        achieved = int8_ops / (pm * 1e-3) * 1e-12
        peak = 2.0 * bf16_peak
        roof = {"bound": "tensor", "kernel": "bgp_potrf_aug: oz_mma_persistent_kernel (tcgen05.mma kind::i8 / UTCIMMA, TMEM accumulators, cp.async.bulk "
                                             "pipeline) trailing updates incl. the M query rows + DMMA panel/leaf kernels",
                "achieved": achieved, "peak": peak, "unit": "TOP/s (int8, 36 int8 ops per fp64 flop of the Ozaki scheme)",
                "frac": achieved / peak, "peak_source": "2 x " + bf16_src + " (no int8 entry; kind::i8 issues at twice the kind::f16 rate)",
                "fp64_equivalent_tflops": fp64_equiv, "fp64_dmma_peak_tflops": FP64_DMMA_PEAK_TFLOPS,
                "fp64_equivalent_over_dmma_peak": fp64_equiv / FP64_DMMA_PEAK_TFLOPS,
                "algorithmic_flop_per_launch": n ** 3 / 3.0, "traffic": TRAFFIC_BYTES_PER_POTRF.get(n)}
    else:
        roof = {"bound": "tensor", "kernel": "bgp_potrf (gemm_nt_kernel DMMA.8x8x4 trailing updates + leaf/panel kernels)",
                "achieved": fp64_equiv, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": fp64_equiv / FP64_DMMA_PEAK_TFLOPS,
                "peak_source": "measured FP64 DMMA microbenchmark on this pool (profiles/fp64_peak_r01.txt); "
                               "MEASURED_PEAKS.json records no fp64 peak (tcgen05 has no f64 kind)",
                "algorithmic_flop_per_launch": n ** 3 / 3.0, "traffic": TRAFFIC_BYTES_PER_POTRF.get(n)}
    line = {
        "metric": "exact_gp_fit_predict_gflops", "value": value, "unit": "GF/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "n": n, "m_query": M_QUERY, "kernel": "wiener+rbf_ard",
                   "l2_policy": "inputs_exceed_l2 (K is %.1f GB per GPU, rebuilt every step)" % (8.0 * n * n / 1e9),
                   "fit_predict_seconds": sec, "potrf_ms": pm, "trailing_update_path": "int8_tcgen05_ozaki" if eng.ozaki else "fp64_dmma",
                   "residual_Kalpha_minus_y_over_y": resid, "lml": lml_chk, "predictions_finite_and_positive": finite},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "GF/s", "seconds": sec_e2e,
                "h2d_bytes_per_step": int(xh.numel() + yh.numel() + xqh.numel()) * 8, "d2h_bytes_per_step": 2 * M_QUERY * 8},
        "gpu_launches": int(launches),
        "roofline": roof,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=40000, help="training points per GP")
    ap.add_argument("--workload", default="per_gpu", choices=["per_gpu", "sharded"])
    ap.add_argument("--ref-n", type=int, default=10000, help="sample size of the CPU arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true", help="sharded workload: also report the matrix-free residual |K alpha - y|/|y|")
    ap.add_argument("--phases", action="store_true", help="sharded workload: extra synchronised pass reporting seconds per phase")
    ap.add_argument("--nb", type=int, default=1024, help="stripe height of the sharded workload")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, args.gpus)
    else:
        run_gpu(args, args.gpus)


if __name__ == "__main__":
    main()
