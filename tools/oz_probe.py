"""Development probe for the int8/tcgen05 trailing-update kernel (csrc/ozaki.cu) -- NOT the bench.

  python tools/oz_probe.py timeline [K]        per-tile clock64 timeline of CTA 0 (bgp_debug_oz_timeline)
  python tools/oz_probe.py raster              event timings of the SYRK 16384^2 for raster / persistence variants
  python tools/oz_probe.py one GROUP K [TPC]   three launches of one configuration (to be wrapped in ncu)
Prints JSON lines."""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
from battgp_b200 import engine as E

dev = torch.device("cuda:0")
eng = E.get_engine(dev)


def ev(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def syrk_setup(n, k):
    torch.manual_seed(0)
    A = torch.randn(n, k, dtype=torch.float64, device=dev)
    Cm = torch.zeros(n, n, dtype=torch.float64, device=dev)
    buf = eng.oz_slice(A)
    return A, Cm, buf


def syrk(buf, n, k, Cm, tpc=0):
    # bgp_oz_gemm runs fully persistent; tiles-per-CTA variants go through the ctx knob used inside bgp_potrf
    eng.oz_gemm(buf, n, 0, buf, n, 0, Cm, k, alpha=-1.0, tri=True)


what = sys.argv[1] if len(sys.argv) > 1 else "timeline"
if what == "timeline":
    ks = [int(a) for a in sys.argv[2:]] or [2048, 1024]
    for cl, k in [(c, k) for k in ks for c in (1, 2)]:
        eng.set("oz_cluster", cl)
        n = 16384
        A, Cm, buf = syrk_setup(n, k)
        cap = 120
        dbg = torch.zeros(cap * 16, dtype=torch.int64, device=dev)
        syrk(buf, n, k, Cm)
        torch.cuda.synchronize()
        eng.L.bgp_debug_oz_timeline(eng.h, C.c_void_p(dbg.data_ptr()), cap)
        syrk(buf, n, k, Cm)
        torch.cuda.synchronize()
        eng.L.bgp_debug_oz_timeline(eng.h, C.c_void_p(0), 0)
        t = dbg.view(cap, 16).cpu().numpy()
        rows = []
        for i in range(2, min(cap, 60)):
            if t[i, 1] == 0 or t[i + 1, 1] == 0:
                break
            rows.append({"tile": i, "period": int(t[i + 1, 1] - t[i, 1]), "mma_wait_tempty": int(t[i, 1] - t[i, 0]),
                         "mma_first_full_wait": int(t[i, 2] - t[i, 1]), "mma_issue_loop": int(t[i, 3] - t[i, 1]),
                         "epi_prefetch_to_tfull": int(t[i, 5] - t[i, 4]), "epi_drain": int(t[i, 6] - t[i, 5]),
                         "epi_store_after_release": int(t[i, 7] - t[i, 6]),
                         "epi_h0_loads": int(t[i, 10] - t[i, 6]), "epi_h0_stores": int(t[i, 11] - t[i, 10]),
                         "bubble_tfull_to_next_mma": int(t[i + 1, 1] - t[i, 5]),
                         "epi_prev_store_end_to_tfull": int(t[i, 5] - t[i - 1, 7]),
                         "prod_tile": int(t[i, 9] - t[i, 8])})
        import statistics as st
        keys = [k2 for k2 in rows[0] if k2 != "tile"]
        print(json.dumps({"op": "oz_timeline", "n": n, "K": k, "cluster": cl, "tiles": len(rows),
                          "median_cycles": {k2: int(st.median(r[k2] for r in rows)) for k2 in keys}}), flush=True)
        for r in rows[:3]:
            print(json.dumps(r), flush=True)
        del A, Cm, buf
elif what == "raster":
    n = 16384
    for k in (2048, 1024):
        A, Cm, buf = syrk_setup(n, k)
        for cl in (1, 2, 4):
            eng.set("oz_cluster", cl)
            for grp in ((8, 16) if cl == 1 else (8,)):
                eng.set("oz_group", grp)
                ms = ev(lambda: syrk(buf, n, k, Cm), reps=5)
                print(json.dumps({"op": "oz_syrk", "n": n, "K": k, "cluster": cl, "group": grp, "ms": round(ms, 4),
                                  "tflops_equiv": round(n * n * k / ms * 1e-9, 2)}), flush=True)
        eng.set("oz_group", 8)
        del A, Cm, buf
elif what == "l2hint":
    n = 16384
    eng.set("oz_cluster", 1)
    for k in (2048, 1024):
        A, Cm, buf = syrk_setup(n, k)
        for hint in (0, 1, 2, 3):
            eng.set("oz_l2hint", hint)
            ms = ev(lambda: syrk(buf, n, k, Cm), reps=5)
            print(json.dumps({"op": "oz_syrk_l2hint", "n": n, "K": k, "l2hint": hint, "ms": round(ms, 4),
                              "tflops_equiv": round(n * n * k / ms * 1e-9, 2)}), flush=True)
        eng.set("oz_l2hint", 0)
        del A, Cm, buf
elif what == "collector":
    # A-collector reuse on/off: timing + bit-identity of the result (integer products are exact, so the two forms must agree exactly)
    n = 16384
    eng.set("oz_cluster", 1)
    for k in (2048, 1024, 512):
        A, Cm, buf = syrk_setup(n, k)
        res = {}
        for on in (0, 1):
            eng.set("oz_collector", on)
            Cm.zero_()
            syrk(buf, n, k, Cm)
            torch.cuda.synchronize()
            res[on] = Cm.clone()
            ms = ev(lambda: syrk(buf, n, k, Cm), reps=5)
            print(json.dumps({"op": "oz_syrk_collector", "n": n, "K": k, "collector": on, "ms": round(ms, 4),
                              "tflops_equiv": round(n * n * k / ms * 1e-9, 2)}), flush=True)
        ref = -(A[:512] @ A[:512].T)
        print(json.dumps({"op": "oz_collector_identity", "K": k, "bit_identical": bool(torch.equal(res[0], res[1])),
                          "max_abs_diff": float((res[0] - res[1]).abs().max()),
                          "max_err_vs_fp64_matmul_512": float((torch.tril(res[1][:512, :512]) - torch.tril(ref)).abs().max())}), flush=True)
        eng.set("oz_collector", 0)
        del A, Cm, buf, res
elif what == "bound":
    # what bounds the k-loop: the same SYRK with operand loads switched off after the first stages (gemm_cfg 7: MMA + smem reads
    # only, results are garbage), with the A-collector on/off, and the CTA-0 timeline of each
    n, k = 16384, 2048
    eng.set("oz_cluster", 1)
    A, Cm, buf = syrk_setup(n, k)
    cap = 120
    ref = {}
    for name, cfg, order in (("loads_on_order0", 0, 0), ("loads_on_order2", 0, 2), ("loads_on_order2_relay", 0, 12), ("loads_on_order2_nofence", 0, 102),
                             ("loads_on_order2_relay_nofence", 0, 112), ("loads_off_order2", 7, 2), ("loads_off_order2_relay_nofence", 7, 112)):
        eng.set("oz_kfence", 0 if order >= 100 else 1)
        order = order % 100
        eng.set("oz_relay", 1 if order >= 10 else 0)
        order = order % 10
        eng.set("gemm_cfg", cfg); eng.set("oz_order", order)
        if cfg == 0:
            Cm.zero_(); syrk(buf, n, k, Cm); torch.cuda.synchronize(); ref[order] = Cm.clone()
        ms = ev(lambda: syrk(buf, n, k, Cm), reps=5)
        dbg = torch.zeros(cap * 16, dtype=torch.int64, device=dev)
        eng.L.bgp_debug_oz_timeline(eng.h, C.c_void_p(dbg.data_ptr()), cap)
        syrk(buf, n, k, Cm)
        torch.cuda.synchronize()
        eng.L.bgp_debug_oz_timeline(eng.h, C.c_void_p(0), 0)
        t = dbg.view(cap, 16).cpu().numpy()
        import statistics as st
        loops = [int(t[i, 3] - t[i, 1]) for i in range(2, 50) if t[i + 1, 1] != 0]
        per = [int(t[i + 1, 1] - t[i, 1]) for i in range(2, 50) if t[i + 1, 1] != 0]
        print(json.dumps({"op": "oz_bound", "case": name, "n": n, "K": k, "ms": round(ms, 4), "tflops_equiv": round(n * n * k / ms * 1e-9, 2),
                          "mma_loop_cycles_per_kblock": round(st.median(loops) / (k / 64), 1), "tile_period_cycles": int(st.median(per))}), flush=True)
    print(json.dumps({"op": "oz_order_identity", "bit_identical": bool(torch.equal(ref[0], ref[2]))}), flush=True)
    for kk in (1024, 512):
        A2, C2, b2 = syrk_setup(n, kk)
        eng.set("gemm_cfg", 0)
        for order in (2, 12):
            eng.set("oz_order", order % 10); eng.set("oz_relay", 1 if order >= 10 else 0)
            ms = ev(lambda: syrk(b2, n, kk, C2), reps=5)
            print(json.dumps({"op": "oz_syrk_order", "n": n, "K": kk, "order": order % 10, "relay": order >= 10, "ms": round(ms, 4), "tflops_equiv": round(n * n * kk / ms * 1e-9, 2)}), flush=True)
        del A2, C2, b2
    eng.set("oz_order", 2); eng.set("oz_relay", 1); eng.set("oz_kfence", 1)
    eng.set("gemm_cfg", 0); eng.set("oz_collector", 0)
elif what == "epi":
    # which part of the epilogue slows the concurrent MMA loop down?  (results are garbage for dbg_epi != 0)
    n = 16384
    eng.set("oz_cluster", 1)
    cap = 120
    import statistics as st
    for k in (2048, 512):
        A, Cm, buf = syrk_setup(n, k)
        for name, de in (("full_epilogue", 0), ("no_store_phase", 1), ("smem_transpose_only", 2), ("global_traffic_only", 3)):
            eng.set("oz_dbg_epi", de)
            ms = ev(lambda: syrk(buf, n, k, Cm), reps=5)
            dbg = torch.zeros(cap * 16, dtype=torch.int64, device=dev)
            eng.L.bgp_debug_oz_timeline(eng.h, C.c_void_p(dbg.data_ptr()), cap)
            syrk(buf, n, k, Cm)
            torch.cuda.synchronize()
            eng.L.bgp_debug_oz_timeline(eng.h, C.c_void_p(0), 0)
            t = dbg.view(cap, 16).cpu().numpy()
            loops = [int(t[i, 3] - t[i, 1]) for i in range(2, 50) if t[i + 1, 1] != 0]
            per = [int(t[i + 1, 1] - t[i, 1]) for i in range(2, 50) if t[i + 1, 1] != 0]
            print(json.dumps({"op": "oz_epi", "case": name, "n": n, "K": k, "ms": round(ms, 4), "tflops_equiv": round(n * n * k / ms * 1e-9, 2),
                              "mma_loop_cycles_per_kblock": round(st.median(loops) / (k / 64), 1), "tile_period_cycles": int(st.median(per))}), flush=True)
        eng.set("oz_dbg_epi", 0)
        del A, Cm, buf
elif what == "backoff":
    n = 16384
    eng.set("oz_cluster", 1)
    cap = 120
    import statistics as st
    for k in (2048, 512):
        A, Cm, buf = syrk_setup(n, k)
        ref = None
        for rep in range(2):
            for bo in (0, 1):
                eng.set("oz_backoff", bo)
                Cm.zero_(); syrk(buf, n, k, Cm); torch.cuda.synchronize()
                if ref is None: ref = Cm.clone()
                same = bool(torch.equal(ref, Cm))
                ms = ev(lambda: syrk(buf, n, k, Cm), reps=5)
                dbg = torch.zeros(cap * 16, dtype=torch.int64, device=dev)
                eng.L.bgp_debug_oz_timeline(eng.h, C.c_void_p(dbg.data_ptr()), cap)
                syrk(buf, n, k, Cm)
                torch.cuda.synchronize()
                eng.L.bgp_debug_oz_timeline(eng.h, C.c_void_p(0), 0)
                t = dbg.view(cap, 16).cpu().numpy()
                loops = [int(t[i, 3] - t[i, 1]) for i in range(2, 50) if t[i + 1, 1] != 0]
                per = [int(t[i + 1, 1] - t[i, 1]) for i in range(2, 50) if t[i + 1, 1] != 0]
                print(json.dumps({"op": "oz_backoff", "backoff": bo, "n": n, "K": k, "ms": round(ms, 4), "tflops_equiv": round(n * n * k / ms * 1e-9, 2),
                                  "mma_loop_cycles_per_kblock": round(st.median(loops) / (k / 64), 1), "tile_period_cycles": int(st.median(per)),
                                  "bit_identical": same}), flush=True)
        eng.set("oz_backoff", 0)
        del A, Cm, buf
elif what == "one":
    grp, k = int(sys.argv[2]), int(sys.argv[3])
    n = 16384
    eng.set("oz_cluster", 1)
    eng.set("oz_group", grp)
    if len(sys.argv) > 4:
        eng.set("oz_l2hint", int(sys.argv[4]))
    A, Cm, buf = syrk_setup(n, k)
    for _ in range(3):
        syrk(buf, n, k, Cm)
    torch.cuda.synchronize()
    print(json.dumps({"op": "oz_one", "group": grp, "K": k}))
