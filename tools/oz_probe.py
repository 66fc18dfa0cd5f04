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
