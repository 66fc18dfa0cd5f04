#!/usr/bin/env python
"""bench.py -- exact-GP fit+predict throughput on B200 (BASELINE.json metric), see DESIGN.md "Measurement".

A *step* is one full pass of the hot path over one synthetic 8s1p data set (SURVEY.md 8d generator):
    build K (Wiener + RBF-ARD + noise)  ->  Cholesky  ->  alpha  ->  LML  ->  predict mean/var at M=300 queries.
N=1 GPU   : BASELINE.json configs[1]  (full_gp, N=40 000, D=3(+t), 1xB200).
N>1 GPUs  : configs[3]  (one independent N=40 000 GP per GPU, no collective; weak scaling) -- the reference's own
            process-per-GPU layout (gp_runner.py:246-298) -- is the line's `value`; the SAME line carries a "sharded" object:
            configs[4], ONE N=200 000 GP block-row-sharded over all ranks with NCCL panel exchanges (strong scaling), with its
            seconds per step, per-phase split, NCCL bytes and parity numbers (oracle at n=5000, 1-GPU engine at N=40 000,
            matrix-free residual at full size).  `--workload sharded` prints configs[4] as the primary line instead.
--workload train : configs[2] (Matern-5/2 + Periodic, N=80 000, LML + analytic-gradient passes; a step = one pass).

value  = algorithmic GFLOP/s (N^3/3 + N^2 M + 2 N^2 per GP) with X, y resident in HBM.
e2e    = the same metric through the public API with HOST (pinned) inputs and host outputs inside the timed region.
--impl reference times the reference's CPU path for the SAME config on the box's host cores: the torch fp64 restatement of
what GPyTorch executes under max_cholesky_size(N+1) (oracle/torch_ref.py; BASELINE.md section 3 -- GPyTorch itself is not
installable here, DESIGN.md "Reference arm").  One N=40 000 step costs ~1 min of CPU, so the arm runs as many of the
requested steps as fit `--ref-budget-s` (at least one) and reports `steps_run`.
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
OZ_PAIRS = 28                  # int8 digit-plane products per fp64 product (csrc/ozaki.cu: 7 planes, pairs with s+t <= 6)
METRIC = "exact_gp_fit_predict_gflops"


def algorithmic_flops(n: int, m: int = M_QUERY) -> float:
    return n ** 3 / 3.0 + float(n) ** 2 * m + 2.0 * float(n) ** 2


def train_pass_flops(n: int) -> float:
    """LML + gradient pass: POTRF N^3/3 + POTRI 2N^3/3 + the O(N^2) sweeps."""
    return float(n) ** 3 + 4.0 * float(n) ** 2


def int8_peak() -> dict:
    """tcgen05 kind::i8 peak measured on this pool (tools/microbench/i8_peak.cu -> profiles/i8_peak_r02.jsonl);
    MEASURED_PEAKS.json records bf16 only.  Fallback: 2 x its bf16 figure, labelled as derived."""
    try:
        rows = [json.loads(l) for l in open(os.path.join(ROOT, "profiles", "i8_peak_r02.jsonl")) if l.strip().startswith("{")]
        r = next(x for x in rows if x.get("mode") == "N256")
        return {"sustained": float(r["sustained_TOPs"]), "burst": float(r["burst_TOPs"]),
                "source": "profiles/i8_peak_r02.jsonl (tools/microbench/i8_peak.cu, M=128 N=256 K=32 tcgen05.mma kind::i8 "
                          "back to back from shared memory; sustained = 1 s under the 1 kW cap, burst = best 2 ms)"}
    except Exception:
        try:
            bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
            src = "DERIVED: 2 x MEASURED_PEAKS.json bf16_tflops (profiles/i8_peak_r02.jsonl missing)"
        except Exception:
            bf16, src = 1590.0, "DERIVED: 2 x fallback 1.59 PFLOP/s (B200_PROFILING.md)"
        return {"sustained": 2.0 * bf16, "burst": 2.0 * bf16, "source": src}


def ncu_traffic(kernel: str):
    """dram bytes per launch of the dominant kernel from this round's committed `ncu --set full` capture (None when absent)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_r02_traffic.json")))
        return d.get(kernel)
    except Exception:
        return None


def step_roofline(n: int, sec: float, world: int, ozaki: bool) -> dict:
    """Roofline of a whole sharded step (N^3/3 of N^3/3 + N^2 M + 2 N^2 flop is the factorisation), per GPU."""
    fp64_equiv = (n ** 3 / 3.0) / sec * 1e-12 / world
    if not ozaki:
        return {"bound": "tensor", "achieved": fp64_equiv, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s per GPU (whole step)",
                "frac": fp64_equiv / FP64_DMMA_PEAK_TFLOPS, "traffic": None}
    pk = int8_peak()
    ach = OZ_PAIRS * fp64_equiv
    return {"bound": "tensor", "achieved": ach, "peak": pk["sustained"],
            "unit": f"TOP/s int8 per GPU (whole step; {OZ_PAIRS} int8 ops per fp64 flop)",
            "frac": ach / pk["sustained"], "peak_source": pk["source"], "fp64_equivalent_tflops_per_gpu": fp64_equiv,
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


# ---------------------------------------------------------------------------------------------------- workload naming
def workload_config(args, n_gpus: int) -> dict:
    """The `config` object: what is computed, nothing measured -- identical in the b200 and the reference arm."""
    n = args.n
    if args.workload == "train":
        return {"workload": f"full_gp Matern-5/2-ARD(I,SOC,T) + Periodic(t) N={n}, LML + analytic-gradient passes "
                            f"(hyper-parameter optimisation step), 1xB200 (BASELINE configs[2])",
                "n": n, "kernel": "matern52_ard+periodic", "l2_policy": "inputs_exceed_l2 (K and K^-1 are %.1f GB each, rebuilt every pass)" % (8.0 * n * n / 1e9)}
    if args.workload == "sharded":
        wl = f"full_gp Wiener+RBF-ARD N={n} block-row-sharded Cholesky over {n_gpus} GPU(s) (BASELINE configs[4])"
    elif n_gpus == 1:
        wl = f"full_gp Wiener+RBF-ARD N={n} D=3(+t) fit+predict on 1xB200 (BASELINE configs[1])"
    else:
        wl = f"8s1p per-cell batch: {n_gpus} independent full_gp N={n}, one per GPU, no collective (BASELINE configs[3])"
    return {"workload": wl, "n": n, "m_query": M_QUERY, "kernel": "wiener+rbf_ard",
            "l2_policy": "inputs_exceed_l2 (K is %.1f GB per GP, rebuilt every step)" % (8.0 * n * n / 1e9)}


def _all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 for multi-rank launches; the CPU arm must still get every host core."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(n)
        return int(torch.get_num_threads())
    except Exception:
        return n


def _cpu_info() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ---------------------------------------------------------------------------------------------------- reference arm
def _cpu_reference_steps(n: int, max_steps: int, budget_s: float, warm: bool = True):
    """Runs the torch fp64 reference path (oracle/torch_ref.py) at size n: returns (seconds per step, steps_run, phases)."""
    from oracle import gp_oracle as orc
    from oracle import torch_ref
    if warm:                                    # thread pools / page faults only; untimed
        xw, yw = orc.synth_field_data(min(n, 2000), seed=1)
        torch_ref.fit_predict(xw, yw, orc.query_grid(xw), noise=NOISE)
    x, y = orc.synth_field_data(n, seed=0)
    xq = orc.query_grid(x)
    times, phases = [], None
    t_start = time.perf_counter()
    while len(times) < max(1, max_steps):
        t0 = time.perf_counter()
        r = torch_ref.fit_predict(x, y, xq, noise=NOISE)
        times.append(time.perf_counter() - t0)
        phases = r["seconds"]
        del r
        if time.perf_counter() - t_start + times[-1] > budget_s:
            break
    return statistics.mean(times), len(times), phases


def _cpu_train_pass(n: int):
    """One LML + gradient pass of configs[2] through the numpy oracle (seconds)."""
    from oracle import gp_oracle as orc
    x, y = orc.synth_field_data(n, seed=0)
    t0 = time.perf_counter()
    orc.lml_grad(orc.matern_periodic_spec(), x, y, NOISE)
    return time.perf_counter() - t0


def run_reference(args, n_gpus: int):
    """CPU arm (rank 0 only): the reference's path for the same config on the host cores."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = _all_host_threads()
    cfg = workload_config(args, n_gpus)
    if args.workload == "train":
        ns = args.ref_n if args.ref_n > 0 else min(args.n, 6000)
        dt_s = _cpu_train_pass(ns)
        dt = dt_s * (args.n / ns) ** 3
        gf = train_pass_flops(args.n) / dt * 1e-9
        sample = (f"one LML+gradient pass of the numpy/LAPACK oracle at N={ns} ({dt_s:.1f} s), EXTRAPOLATED x (N/{ns})^3 to N={args.n} "
                  f"(K and K^-1 at N={args.n} do not fit the time budget on CPU)")
        steps_run, phases, kind = 1, None, "port"
    else:
        # every BASELINE config of this path is built from N=40 000 GPs except configs[4] (N=200 000: 320 GB, hours of CPU --
        # measure at 40k and label the rate as extrapolated)
        ns = args.ref_n if args.ref_n > 0 else min(args.n, 40000)
        dt_s, steps_run, phases = _cpu_reference_steps(ns, args.steps, args.ref_budget_s)
        gf = algorithmic_flops(ns) / dt_s * 1e-9
        dt = algorithmic_flops(args.n) / (gf * 1e9)
        sample = (f"{steps_run} fit+predict step(s) at N={ns} (same generator/hyper-parameters), torch fp64 on {threads} threads: "
                  f"vectorised build -> linalg.cholesky_ex -> cholesky_solve -> solve_triangular (M={M_QUERY})")
        if ns != args.n:
            sample += f"; GF/s is the measured N={ns} rate, ms_per_step EXTRAPOLATED to N={args.n} at that rate"
        if n_gpus > 1:
            sample += f"; one CPU job (the {n_gpus} GPs of the batch would run one after another at this rate)"
        kind = "port"
    line = {"impl": "reference", "metric": METRIC if args.workload != "train" else "exact_gp_lml_grad_pass_gflops",
            "value": gf, "unit": "GF/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "steps_run": steps_run, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "detail": {"sample_n": ns, "phase_seconds": phases, "cpu": _cpu_info(), "os_cpu_count": os.cpu_count(),
                       "time_budget_s": args.ref_budget_s,
                       "note": "steps_run of the requested steps were executed inside the time budget; ms_per_step is their mean"},
            "cpu_baseline": {"value": gf, "unit": "GF/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": gf, "unit": "GF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(args):
    threads = _all_host_threads()
    ns = args.ref_n if args.ref_n > 0 else min(args.n, 40000)
    dt, steps_run, phases = _cpu_reference_steps(ns, 1, 0.0)
    return {"value": algorithmic_flops(ns) / dt * 1e-9, "unit": "GF/s", "cores": threads, "kind": "port",
            "sample": f"one fit+predict at N={ns} (same generator/hyper-parameters), {dt:.1f} s of torch fp64 "
                      f"(build {phases['build']:.1f} s, cholesky_ex {phases['cholesky']:.1f} s, predict {phases['predict']:.1f} s) on {_cpu_info()}"}


# ---------------------------------------------------------------------------------------------------- GPU arm
def _init_dist(n_gpus: int):
    import torch
    import torch.distributed as dist
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
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=600))
    return rank, local_rank, world, dev


def _oracle_checker(x, y, xq):
    """CPU oracle as the CHECKER of the sharded parity numbers (never on the measured path)."""
    from oracle import gp_oracle as orc
    f = orc.fit(orc.battgp_spec(), x, y, NOISE)
    m, v = orc.predict(orc.battgp_spec(), x, f, xq)
    return m, v, f.lml


def run_gpu(args, n_gpus: int):
    import torch
    import torch.distributed as dist
    from battgp_b200 import engine as E
    from battgp_b200.synth import query_grid, synth_field_data

    rank, local_rank, world, dev = _init_dist(n_gpus)
    if args.workload == "sharded":
        from battgp_b200 import sharded
        line = sharded.bench(args, rank, world, dev, checker=_oracle_checker)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "train":
        return run_train(args, rank, world, dev)

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
    eng.kernel_profile(True)                      # CUDA events around every trailing-update launch (bgp_ctx_oz_profile)
    ms_total = timed(step_device, args.steps, collect_potrf=True)
    kprof = eng.kernel_profile(False)
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

    sharded_obj = None
    if world > 1 and not args.no_sharded:
        # configs[4] in the same run: free this rank's per-GPU buffers first (N=200k on 2 GPUs needs 80 GB per rank)
        del K, xd, yd, xqd
        eng.release_workspace()
        torch.cuda.empty_cache()
        from battgp_b200 import sharded
        try:
            sharded_obj = sharded.bench_object(args, rank, world, dev, checker=_oracle_checker)
        except Exception as e:          # the per-GPU line above stays valid; the failure is reported, not hidden
            sharded_obj = {"error": f"{type(e).__name__}: {e}"[:800]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pm = statistics.mean(potrf_ms)
    fp64_equiv = n ** 3 / 3.0 / (pm * 1e-3) * 1e-12
    if eng.ozaki:
        # dominant kernel: the int8/tcgen05 trailing updates (csrc/ozaki.cu oz_mma_kernel), timed live with CUDA events on the
        # stream they are launched on, inside the timed steps; algorithmic flop = 2 x (tile area actually computed) x K
        pk = int8_peak()
        k_ms, k_flop, k_n = kprof["ms"], kprof["flop"], kprof["launches"]
        k_fp64 = k_flop / (k_ms * 1e-3) * 1e-12 if k_ms > 0 else 0.0
        achieved = OZ_PAIRS * k_fp64
        roof = {"bound": "tensor",
                "kernel": "oz_mma_kernel (csrc/ozaki.cu): trailing updates A22 -= L21 L21^T as 28 exact int8 digit-plane products per "
                          "fp64 product on tcgen05.mma kind::i8 (UTCIMMA), accumulators in TMEM, cp.async.bulk operand pipeline",
                "achieved": achieved, "peak": pk["sustained"], "unit": f"TOP/s (int8; {OZ_PAIRS} int8 ops per fp64 flop)",
                "frac": achieved / pk["sustained"], "peak_source": pk["source"], "peak_burst": pk["burst"],
                "frac_of_burst_peak": achieved / pk["burst"],
                "launches_timed": k_n, "avg_launch_ms": k_ms / max(k_n, 1),
                "algorithmic_flop_per_launch": k_flop / max(k_n, 1), "kernel_fp64_equivalent_tflops": k_fp64,
                "kernel_share_of_step": (k_ms / args.steps) / (sec * 1e3),
                "note": "launch durations are CUDA-event times on the launching stream while the look-ahead panel stream shares the SMs",
                "potrf": {"ms": pm, "fp64_equivalent_tflops": fp64_equiv, "fp64_dmma_peak_tflops": FP64_DMMA_PEAK_TFLOPS,
                          "fp64_equivalent_over_dmma_peak": fp64_equiv / FP64_DMMA_PEAK_TFLOPS,
                          "int8_tops": OZ_PAIRS * fp64_equiv, "frac_of_int8_peak": OZ_PAIRS * fp64_equiv / pk["sustained"]},
                "traffic": (ncu_traffic("oz_mma_kernel") or {}).get("dram_bytes_per_launch"),
                "traffic_source": (ncu_traffic("oz_mma_kernel") or {}).get("source", "no ncu capture of this round committed")}
    else:
        roof = {"bound": "tensor", "kernel": "bgp_potrf (gemm_nt_kernel DMMA.8x8x4 trailing updates + leaf/panel kernels)",
                "achieved": fp64_equiv, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": fp64_equiv / FP64_DMMA_PEAK_TFLOPS,
                "peak_source": "measured FP64 DMMA microbenchmark on this pool (profiles/fp64_peak_r01.txt); "
                               "MEASURED_PEAKS.json records no fp64 peak (tcgen05 has no f64 kind)",
                "algorithmic_flop_per_launch": n ** 3 / 3.0, "traffic": None}
    line = {
        "metric": METRIC, "value": value, "unit": "GF/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "detail": {"fit_predict_seconds": sec, "potrf_ms": pm, "trailing_update_path": "int8_tcgen05_ozaki_7x8bit" if eng.ozaki else "fp64_dmma",
                   "residual_Kalpha_minus_y_over_y": resid, "lml": lml_chk, "predictions_finite_and_positive": finite},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "GF/s", "seconds": sec_e2e,
                "h2d_bytes_per_step": int(xh.numel() + yh.numel() + xqh.numel()) * 8, "d2h_bytes_per_step": 2 * M_QUERY * 8},
        "gpu_launches": int(launches),
        "roofline": roof,
    }
    if sharded_obj is not None:
        line["sharded"] = sharded_obj
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train(args, rank: int, world: int, dev):
    """BASELINE configs[2]: Matern-5/2-ARD + Periodic, LML + analytic gradient passes at N (default 80 000) on one GPU.
    A step = build K -> Cholesky -> alpha -> LML -> K^-1 (POTRI) -> fused gradient reduction (6+ hyper-parameters)."""
    import numpy as np
    import torch
    from battgp_b200 import engine as E
    from battgp_b200.synth import synth_field_data
    if world != 1:
        raise SystemExit("--workload train runs on one GPU")
    eng = E.get_engine(dev)
    n = args.n
    spec = E.matern_periodic_spec()
    x_np, y_np = synth_field_data(n, seed=0)
    xd, yd = torch.tensor(x_np, device=dev), torch.tensor(y_np, device=dev)
    xh, yh = torch.tensor(x_np).pin_memory(), torch.tensor(y_np).pin_memory()
    K = E.alloc_matrix(n, n, dev)
    W = E.alloc_matrix(n, n, dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    phase = {"fit": [], "potri": [], "grad": []}

    def one_pass(x, y, timed_phases=False):
        if timed_phases:
            ev[0].record()
        st = E.fit(spec, x, y, NOISE, K_out=K)
        if timed_phases:
            ev[1].record()
        eng.potri(st.L, st.dinv, W)
        if timed_phases:
            ev[2].record()
        g = eng.lml_grad(spec, NOISE, x, st.L, st.alpha)
        if timed_phases:
            ev[3].record()
            torch.cuda.synchronize(dev)
            phase["fit"].append(ev[0].elapsed_time(ev[1])); phase["potri"].append(ev[1].elapsed_time(ev[2]))
            phase["grad"].append(ev[2].elapsed_time(ev[3]))
        return st.lml, g

    def timed(fn, steps):
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1)

    for _ in range(args.warmup):
        one_pass(xd, yd)
    sampler = ClockSampler(dev.index)
    sampler.start()
    l0 = eng.launches
    ms_total = timed(lambda: one_pass(xd, yd, True), args.steps)
    launches = eng.launches - l0
    clocks = sampler.stop()

    def pass_e2e():
        lml, g = one_pass(xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True))
        return lml, g.cpu()
    pass_e2e()
    ms_e2e = timed(pass_e2e, args.steps)
    lml_full, g_full = one_pass(xd, yd)
    # parity of loss and gradient against the oracle at n=1500 in the same run (checker only)
    from oracle import gp_oracle as orc
    nchk = 1500
    xc, yc = synth_field_data(nchk, seed=5)
    st = E.fit(spec, torch.tensor(xc, device=dev), torch.tensor(yc, device=dev), NOISE)
    eng.potri(st.L, st.dinv)
    gc = eng.lml_grad(spec, NOISE, st.x, st.L, st.alpha).cpu().numpy()
    ref = orc.lml_grad(orc.matern_periodic_spec(), xc, yc, NOISE)
    gref = [ref["noise"]]
    for t in ref["terms"]:
        gref += [t["outputscale"], *t["lengthscale"], *t["period"]]
    gref = np.asarray(gref)
    par = {"n": nchk, "lml_rel_err": abs(st.lml - ref["lml"]) / abs(ref["lml"]),
           "grad_max_rel_err": float(np.max(np.abs(gc - gref) / np.maximum(np.abs(gref), 1e-300))),
           "tolerance": {"lml_rel": 1e-9, "grad_rel": 1e-6}, "checker": "oracle/gp_oracle.py lml_grad (parity unpinned by the reference)"}
    par["ok"] = bool(par["lml_rel_err"] < 1e-9 and par["grad_max_rel_err"] < 1e-6)
    sec = ms_total / 1e3 / args.steps
    sec_e2e = ms_e2e / 1e3 / args.steps
    potri_s = statistics.mean(phase["potri"]) * 1e-3
    potri_tf = (2.0 / 3.0) * n ** 3 / potri_s * 1e-12
    pk = int8_peak()
    line = {"metric": "exact_gp_lml_grad_pass_gflops", "value": train_pass_flops(n) / sec * 1e-9, "unit": "GF/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, 1),
            "detail": {"seconds_per_pass": sec, "phase_ms": {k: statistics.mean(v) for k, v in phase.items()}, "lml": lml_full,
                       "grad": [float(v) for v in g_full.cpu()], "parity": par,
                       "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9},
            "clocks": clocks,
            "e2e": {"value": train_pass_flops(n) / sec_e2e * 1e-9, "unit": "GF/s", "seconds": sec_e2e,
                    "h2d_bytes_per_step": int(xh.numel() + yh.numel()) * 8, "d2h_bytes_per_step": int(g_full.numel()) * 8 + 8},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "bgp_potri (K^-1 = L^-T L^-1: trtri + lauum products on the int8/tcgen05 path, K-chunked)",
                         "achieved": OZ_PAIRS * potri_tf, "peak": pk["sustained"], "unit": f"TOP/s (int8; {OZ_PAIRS} int8 ops per fp64 flop)",
                         "frac": OZ_PAIRS * potri_tf / pk["sustained"], "peak_source": pk["source"],
                         "algorithmic_flop_per_launch": (2.0 / 3.0) * n ** 3, "fp64_equivalent_tflops": potri_tf,
                         "fp64_equivalent_over_dmma_peak": potri_tf / FP64_DMMA_PEAK_TFLOPS, "traffic": None}}
    if not args.no_cpu_baseline:
        ns = args.ref_n if args.ref_n > 0 else 4000
        dt = _cpu_train_pass(ns)
        line["cpu_baseline"] = {"value": train_pass_flops(ns) / dt * 1e-9, "unit": "GF/s", "cores": _all_host_threads(), "kind": "port",
                                "sample": f"one LML+gradient pass of the numpy/LAPACK oracle at N={ns}: {dt:.1f} s"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=0, help="training points per GP (default 40000; sharded 200000; train 80000)")
    ap.add_argument("--workload", default="per_gpu", choices=["per_gpu", "sharded", "train"])
    ap.add_argument("--ref-n", type=int, default=0, help="size the CPU arm / cpu_baseline really runs (0 = the workload's, capped at 40000)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="CPU arm: stop starting new steps after this many seconds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="N>1: skip the configs[4] object")
    ap.add_argument("--sharded-n", type=int, default=200000, help="N>1: size of the sharded GP in the 'sharded' object")
    ap.add_argument("--sharded-steps", type=int, default=2)
    ap.add_argument("--verify", action="store_true", help="sharded workload: also report the matrix-free residual |K alpha - y|/|y|")
    ap.add_argument("--phases", action="store_true", help="sharded workload: extra synchronised pass reporting seconds per phase")
    ap.add_argument("--nb", type=int, default=1024, help="stripe height of the sharded workload")
    args = ap.parse_args()
    if args.n <= 0:
        args.n = {"per_gpu": 40000, "sharded": 200000, "train": 80000}[args.workload]
    if args.impl == "reference":
        run_reference(args, args.gpus)
    else:
        run_gpu(args, args.gpus)


if __name__ == "__main__":
    main()
