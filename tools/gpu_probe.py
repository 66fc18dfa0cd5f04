"""Development probe (NOT the bench): time the main kernels in isolation on one B200. Prints JSON lines."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from battgp_b200 import engine as E
from battgp_b200.synth import synth_field_data, query_grid

dev = torch.device("cuda:0")
eng = E.get_engine(dev)

def ev(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

what = sys.argv[1:] or ["gemm", "potrf", "fit"]
if "gemm" in what:
    for (M, N, K) in [(8192, 8192, 8192), (16384, 16384, 1024), (16384, 16384, 512), (16384, 16384, 256), (16384, 16384, 128),
                      (32768, 1024, 1024), (4096, 4096, 4096), (2048, 2048, 2048), (1024, 1024, 1024), (512, 512, 512), (300, 20000, 20000)]:
        A = torch.randn(M, K, dtype=torch.float64, device=dev); B = torch.randn(N, K, dtype=torch.float64, device=dev)
        C = torch.zeros(M, N, dtype=torch.float64, device=dev)
        ms = ev(lambda: eng.gemm_nt(A, B, C, alpha=-1.0, beta=1.0))
        ms_cublas = ev(lambda: torch.addmm(C, A, B.t(), beta=1.0, alpha=-1.0, out=C))
        print(json.dumps({"op": "gemm_nt", "M": M, "N": N, "K": K, "ms": ms, "tflops": 2 * M * N * K / ms * 1e-9,
                          "cublas_ms": ms_cublas, "cublas_tflops": 2 * M * N * K / ms_cublas * 1e-9}), flush=True)
        del A, B, C
    n, k = 16384, 1024
    A = torch.randn(n, k, dtype=torch.float64, device=dev); C = torch.zeros(n, n, dtype=torch.float64, device=dev)
    ms = ev(lambda: eng.gemm_nt(A, A, C, alpha=-1.0, beta=1.0, tri=True))
    print(json.dumps({"op": "syrk_tri", "n": n, "k": k, "ms": ms, "tflops": n * n * k / ms * 1e-9}), flush=True)
    del A, C
if "gemmcfg" in what:
    for cfg in (1, 4, 5, 6):
        eng.set("gemm_cfg", cfg)
        for (M, N, K) in [(8192, 8192, 8192), (16384, 16384, 1024), (16384, 16384, 256)]:
            A = torch.randn(M, K, dtype=torch.float64, device=dev); B = torch.randn(N, K, dtype=torch.float64, device=dev)
            C = torch.zeros(M, N, dtype=torch.float64, device=dev)
            ms = ev(lambda: eng.gemm_nt(A, B, C, alpha=-1.0, beta=1.0))
            print(json.dumps({"op": "gemm_nt", "cfg": cfg, "M": M, "N": N, "K": K, "ms": ms, "tflops": 2 * M * N * K / ms * 1e-9}), flush=True)
            del A, B, C
    eng.set("gemm_cfg", 0)
if "gemmone" in what:
    M, N, K = 8192, 8192, 2048
    A = torch.randn(M, K, dtype=torch.float64, device=dev); B = torch.randn(N, K, dtype=torch.float64, device=dev)
    C = torch.zeros(M, N, dtype=torch.float64, device=dev)
    for _ in range(3): eng.gemm_nt(A, B, C, alpha=-1.0, beta=1.0)
    torch.cuda.synchronize()
if "potrf" in what:
    for n in ((40000,) if 'only40k' in what else (4096, 8192, 16384, 40000)):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev); spec = E.battgp_spec()
        K = E.alloc_matrix(n, n, dev)
        msb = ev(lambda: eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K))
        print(json.dumps({"op": "cov_build_sym", "n": n, "ms": msb, "GBps": 8 * n * (n + 1) / 2 / msb * 1e-6}), flush=True)
        for nb, la in ((1024, 1), (2048, 1), (1024, 0)):
            if n < 16384 and nb != 1024: continue
            eng.set("nb", nb); eng.set("lookahead", la)
            best = 1e30
            for r in range(2):
                eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(json.dumps({"op": "potrf", "n": n, "nb": nb, "lookahead": la, "info": info, "ms": best,
                              "tflops": n ** 3 / 3 / best * 1e-9}), flush=True)
        eng.set("nb", 1024); eng.set("lookahead", 1)
        if n <= 16384:
            eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
            Kc = K.clone()
            ms = ev(lambda: torch.linalg.cholesky_ex(Kc.copy_(K)), reps=2)
            print(json.dumps({"op": "torch_cholesky_ex(+copy)", "n": n, "ms": ms}), flush=True)
            del Kc
        del K
        torch.cuda.empty_cache()
if "fit" in what:
    for n in (8192, 40000):
        x, y = synth_field_data(n, 0)
        xd, yd = torch.tensor(x, device=dev), torch.tensor(y, device=dev)
        xq = torch.tensor(query_grid(x), device=dev)
        K = E.alloc_matrix(n, n, dev)
        for r in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            st = E.fit(E.battgp_spec(), xd, yd, 2.33e-6, K_out=K)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            m, v = E.predict(st, xq)
            torch.cuda.synchronize(); t2 = time.perf_counter()
        # phase split
        t = {}
        t["build"] = ev(lambda: eng.cov_build(st.spec, xd, noise=2.33e-6, symmetric=True, out=K), reps=2)
        info, ld, dinv = eng.potrf(K)
        t["potrs_vec"] = ev(lambda: eng.potrs_vec(K, dinv, yd), reps=2)
        Kq = eng.cov_build(st.spec, xq, xd)
        t["trsm300"] = ev(lambda: eng.trsm_rlt(K, dinv, Kq), reps=2)
        t["cross_build"] = ev(lambda: eng.cov_build(st.spec, xq, xd, out=Kq), reps=2)
        t["tail_mean"] = ev(lambda: eng.predict_tail(Kq=Kq, alpha=st.alpha), reps=2)
        kd = eng.cov_diag(st.spec, xq)
        t["tail_var"] = ev(lambda: eng.predict_tail(V=Kq, kdiag=kd), reps=2)
        t["predict_events"] = ev(lambda: E.predict(st, xq), reps=2)
        print(json.dumps({"op": "fit+predict", "n": n, "fit_s": t1 - t0, "predict_s": t2 - t1, "lml": st.lml,
                          "mean0": float(m[0]), "var0": float(v[0]), "phases_ms": t, "launches": eng.launches}), flush=True)
        del K, st
        torch.cuda.empty_cache()

if "grad" in what:
    for n, mk in ((20000, "battgp"), (40000, "battgp"), (40000, "matern_periodic")):
        x, y = synth_field_data(n, 0)
        xd, yd = torch.tensor(x, device=dev), torch.tensor(y, device=dev)
        spec = E.battgp_spec() if mk == "battgp" else E.matern_periodic_spec()
        K = E.alloc_matrix(n, n, dev); W = E.alloc_matrix(n, n, dev)
        res = {}
        for r in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            st = E.fit(spec, xd, yd, 2.33e-6, K_out=K)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            eng.potri(st.L, st.dinv, W)
            torch.cuda.synchronize(); t2 = time.perf_counter()
            g = eng.lml_grad(spec, 2.33e-6, xd, st.L, st.alpha)
            torch.cuda.synchronize(); t3 = time.perf_counter()
            res = {"fit_s": t1 - t0, "potri_s": t2 - t1, "grad_s": t3 - t2, "pass_s": t3 - t0}
        print(json.dumps({"op": "lml+grad pass", "n": n, "kernel": mk, **res, "potri_tflops": 0.83 * n ** 3 / res["potri_s"] * 1e-12,
                          "grad": [float(v) for v in g.cpu()]}), flush=True)
        del K, W, st
        torch.cuda.empty_cache()
if "syrkone" in what:
    n, k = 16384, 1024
    A = torch.randn(n, k, dtype=torch.float64, device=dev); C = torch.zeros(n, n, dtype=torch.float64, device=dev)
    for _ in range(3): eng.gemm_nt(A, A, C, alpha=-1.0, beta=1.0, tri=True)
    torch.cuda.synchronize()
if "i8" in what:
    torch.manual_seed(0)
    for (M, N, K, tri) in [(128, 64, 64, False), (128, 64, 128, False), (256, 192, 1024, False), (1000, 700, 512, False), (2048, 2048, 1024, True),
                           (16384, 16384, 512, True), (16384, 16384, 1024, True), (16384, 16384, 2048, True), (16384, 16384, 4096, True)]:
        A = torch.randn(M, K, dtype=torch.float64, device=dev) * torch.exp(torch.randn(M, 1, dtype=torch.float64, device=dev) * 3)
        B = A if tri else torch.randn(N, K, dtype=torch.float64, device=dev) * torch.exp(torch.randn(N, 1, dtype=torch.float64, device=dev) * 3)
        C0 = torch.randn(M, N, dtype=torch.float64, device=dev)
        C1 = C0.clone(); C2 = C0.clone()
        eng.gemm_nt(A, B, C1, alpha=-1.0, beta=1.0, tri=tri)
        work = torch.empty(int(eng.L.bgp_gemm_nt_i8_work_bytes(M, N, K)), dtype=torch.uint8, device=dev)
        eng.gemm_nt_i8(A, B, C2, alpha=-1.0, tri=tri, work=work)
        torch.cuda.synchronize()
        scale = (A.abs().amax(1, keepdim=True) * B.abs().amax(1, keepdim=True).T) * K
        err = ((C1 - C2).abs() / scale).max().item()
        rel = ((C1 - C2).norm() / (C1 - C0).norm()).item()
        res = {"op": "gemm_nt_i8", "M": M, "N": N, "K": K, "tri": tri, "max_err_over_rowmax_colmax_K": err, "rel_fro_vs_dmma": rel}
        if M >= 2048:
            C1ref = C0.clone(); eng.gemm_nt(A, B, C1ref, alpha=-1.0, beta=1.0, tri=tri); torch.cuda.synchronize()
            ms = ev(lambda: eng.gemm_nt_i8(A, B, C2, alpha=-1.0, tri=tri, work=work), reps=3)
            ms_d = ev(lambda: eng.gemm_nt(A, B, C1, alpha=-1.0, beta=1.0, tri=tri), reps=3)
            fl = (1 if tri else 2) * M * N * K
            res.update({"ms_i8_incl_slicing": ms, "tflops_equiv": fl / ms * 1e-9, "ms_dmma": ms_d, "tflops_dmma": fl / ms_d * 1e-9})
            for cs in (2, 4):
                eng.set("oz_cluster", cs)
                C3 = C0.clone()
                eng.gemm_nt_i8(A, B, C3, alpha=-1.0, tri=tri, work=work)
                dd = ((C3 - C2b).abs().max() / C2b.abs().max()).item() if False else 0.0
                ms_c = ev(lambda: eng.gemm_nt_i8(A, B, C3, alpha=-1.0, tri=tri, work=work), reps=3)
                C4 = C0.clone(); eng.gemm_nt_i8(A, B, C4, alpha=-1.0, tri=tri, work=work); torch.cuda.synchronize()
                rel = ((C4 - C1ref).norm() / (C1ref - C0).norm()).item()
                res.update({f"ms_cluster{cs}": ms_c, f"tflops_cluster{cs}": fl / ms_c * 1e-9, f"rel_cluster{cs}": rel})
                del C3, C4
            eng.set("oz_cluster", 1)
        print(json.dumps(res), flush=True)
        del A, B, C0, C1, C2, work

if "ozpotrf" in what:
    for n in (8192, 16384, 40000):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev); spec = E.battgp_spec()
        K = E.alloc_matrix(n, n, dev); Kref = E.alloc_matrix(n, n, dev)
        eng.set("ozaki", 0); eng.set("nb", 1024)
        eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=Kref)
        info0, ld0, _ = eng.potrf(Kref)
        for nb, tpc in ((1024, 4), (2048, 4), (1024, 8), (2048, 2), (1024, 1), (2048, 1), (2048, 8)):
            if n <= 2 * nb: continue
            if n < 40000 and tpc != 4: continue
            eng.set("nb", nb); eng.set("ozaki", 1); eng.set("oz_tpc", tpc)
            best = 1e30
            for r in range(2):
                eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            dL = (torch.tril(K) - torch.tril(Kref)).norm() / torch.tril(Kref).norm()
            print(json.dumps({"op": "potrf_ozaki", "n": n, "nb": nb, "tpc": tpc, "info": info, "ms": best, "tflops_equiv": n ** 3 / 3 / best * 1e-9,
                              "rel_L_vs_dmma": float(dL), "logdet_diff": ld - ld0}), flush=True)
        eng.set("ozaki", 1); eng.set("nb", 0); eng.set("oz_tpc", 4)
        del K, Kref
        torch.cuda.empty_cache()
if "ozone" in what:
    n, k = 16384, 2048
    A = torch.randn(n, k, dtype=torch.float64, device=dev); C = torch.zeros(n, n, dtype=torch.float64, device=dev)
    work = torch.empty(int(eng.L.bgp_gemm_nt_i8_work_bytes(n, n, k)), dtype=torch.uint8, device=dev)
    for _ in range(3): eng.gemm_nt_i8(A, A, C, alpha=-1.0, tri=True, work=work)
    torch.cuda.synchronize()
if "leaf" in what:
    for n in (128, 256, 512, 1024, 2048):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev); spec = E.battgp_spec()
        K0 = eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True)
        Ks = [K0.clone() for _ in range(20)]
        torch.cuda.synchronize()
        info_d = torch.full((1,), 2**31 - 1, dtype=torch.int32, device=dev); ld_d = torch.zeros(1, dtype=torch.float64, device=dev)
        eng.potrf_block(K0.clone(), info_d, ld_d); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for Kc in Ks: eng.potrf_block(Kc, info_d, ld_d)
        e1.record(); torch.cuda.synchronize()
        import ctypes
        clk = (ctypes.c_longlong * 8)()
        eng.L.bgp_debug_leaf_clk(clk)
        nl = 21 * ((n + 127) // 128)
        print(json.dumps({"op": "potrf_block(async)", "n": n, "us_per_call": e0.elapsed_time(e1) / 20 * 1e3,
                          "leaf_cycles_per_phase[load,potrf32x4,subst+inv x4,trail x4,storeL,invphase,storedinv]": [int(c / nl) for c in clk][:7]}), flush=True)

if "nbsweep" in what:
    n = 40000
    x, y = synth_field_data(n, 0)
    xd = torch.tensor(x, device=dev); spec = E.battgp_spec()
    K = E.alloc_matrix(n, n, dev)
    for nb, tpc, la in ((1024, 2, 1), (1536, 2, 1), (2048, 2, 1), (2560, 2, 1), (3072, 2, 1), (2048, 1, 1), (2048, 3, 1), (1536, 1, 1), (2048, 0, 0), (1024, 0, 0)):
        eng.set("nb", nb); eng.set("ozaki", 1); eng.set("oz_tpc", tpc); eng.set("lookahead", la)
        best = 1e30
        for r in range(2):
            eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(json.dumps({"op": "potrf_sweep", "n": n, "nb": nb, "tpc": tpc, "lookahead": la, "info": info, "ms": best}), flush=True)
    eng.set("nb", 0); eng.set("oz_tpc", 2); eng.set("lookahead", 1)
if "small" in what:
    import numpy as np
    for n in (1000, 2000, 4000, 8192):
        x, y = synth_field_data(n, 0)
        xd, yd = torch.tensor(x, device=dev), torch.tensor(y, device=dev)
        xq = torch.tensor(query_grid(x), device=dev)
        spec = E.battgp_spec()
        def ours():
            st = E.fit(spec, xd, yd, 2.33e-6, xq=xq)
            return E.predict(st, xq)
        for _ in range(3): ours()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): ours()
        torch.cuda.synchronize(); t_ours = (time.perf_counter() - t0) / 10
        # yardstick: the same math with torch/cuSOLVER library calls (K built by our kernel)
        def lib():
            K = eng.cov_build(spec, xd, xd)
            K.diagonal().add_(2.33e-6)
            L = torch.linalg.cholesky(K)
            alpha = torch.cholesky_solve(yd[:, None], L)
            Kq = eng.cov_build(spec, xq, xd)
            mean = Kq @ alpha
            V = torch.linalg.solve_triangular(L, Kq.T, upper=False)
            var = eng.cov_diag(spec, xq) - (V * V).sum(0)
            return mean, var
        for _ in range(3): lib()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): lib()
        torch.cuda.synchronize(); t_lib = (time.perf_counter() - t0) / 10
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        K = eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True)
        Ks = [K.clone() for _ in range(5)]
        torch.cuda.synchronize(); ev0.record()
        for Kc in Ks: eng.potrf(Kc)
        ev1.record(); torch.cuda.synchronize()
        print(json.dumps({"op": "small_n_fit_predict", "n": n, "ours_ms": t_ours * 1e3, "torch_cusolver_ms": t_lib * 1e3,
                          "our_potrf_ms": ev0.elapsed_time(ev1) / 5}), flush=True)
if "midsweep" in what:
    for n in (6144, 8192, 12288, 16384, 24576):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev); spec = E.battgp_spec()
        K = E.alloc_matrix(n, n, dev)
        for nb in (256, 512, 768, 1024, 1536):
            if n <= 2 * nb: continue
            eng.set("nb", nb); eng.set("ozaki", 1)
            best = 1e30
            for r in range(3):
                eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(json.dumps({"op": "potrf_mid", "n": n, "nb": nb, "info": info, "ms": round(best, 3)}), flush=True)
        eng.set("nb", 0)
        del K
        torch.cuda.empty_cache()
if "sched" in what:
    # panel-schedule sweep: uniform widths ("nb") against the remaining-rows rule (sched_* knobs of bgp_ctx_set)
    spec = E.battgp_spec()
    DEF = {"sched_t1024": 9000, "sched_t2048": 17000, "sched_t4096": 0, "sched_w0": 0, "sched_w1": 0}
    def run(n, xd, K, cfg, reps=2):
        for k, v in DEF.items(): eng.set(k, cfg.get(k, v))
        eng.set("nb", cfg.get("nb", 0)); eng.set("ozaki", 1)
        best = 1e30
        for r in range(reps):
            eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(json.dumps({"op": "potrf_sched", "n": n, "cfg": cfg, "info": info, "logdet": ld, "ms": round(best, 3)}), flush=True)
    sweeps = {
        40000: [{"nb": 2048}, {}, {"sched_t2048": 10000}, {"sched_t2048": 18000}, {"sched_t2048": 22000},
                {"sched_t1024": 6000}, {"sched_t1024": 12000}, {"sched_w0": 512}, {"sched_w0": 1024},
                {"sched_w0": 1024, "sched_t4096": 30000}, {"sched_w0": 1024, "sched_w1": 2048, "sched_t4096": 26000},
                {"sched_w0": 1024, "sched_w1": 2048, "sched_t4096": 22000}, {"sched_w0": 1024, "sched_t2048": 18000, "sched_t1024": 12000}],
        24576: [{"nb": 1024}, {"nb": 2048}, {}, {"sched_t2048": 18000}, {"sched_w0": 1024}, {"sched_w0": 512}],
        16384: [{"nb": 1024}, {"nb": 512}, {}, {"sched_t2048": 0}, {"sched_t2048": 0, "sched_t1024": 12000}, {"sched_w0": 512}, {"sched_w0": 1024}],
        12288: [{"nb": 512}, {"nb": 1024}, {}, {"sched_t1024": 6000}, {"sched_t1024": 12000}],
        8192: [{"nb": 512}, {}, {"sched_t1024": 6000}],
    }
    for n, cfgs in sweeps.items():
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev)
        K = E.alloc_matrix(n, n, dev)
        for cfg in cfgs: run(n, xd, K, cfg)
        del K
    for k, v in DEF.items(): eng.set(k, v)
    eng.set("nb", 0)
if "trace" in what:
    # per-panel timeline of the look-ahead factorisation ("trace" knob; lines go to stderr)
    spec = E.battgp_spec()
    for n in (40000, 16384):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev)
        K = E.alloc_matrix(n, n, dev)
        eng.set("ozaki", 1)
        for tr in (0, 1):
            eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
            torch.cuda.synchronize()
            eng.set("trace", tr)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
            eng.set("trace", 0)
            print(json.dumps({"op": "potrf_trace_total", "n": n, "trace": tr, "ms": e0.elapsed_time(e1)}), flush=True)
        del K
if "pdl" in what:
    # programmatic dependent launch on the chain kernels (leaf + DMMA GEMMs): on/off
    spec = E.battgp_spec()
    for n in (1024, 2048, 4096, 8192, 16384, 40000):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev)
        K = E.alloc_matrix(n, n, dev)
        eng.set("ozaki", 1)
        for pdl in (0, 1, 0, 1):
            eng.set("pdl", pdl)
            best = 1e30
            for r in range(4 if n < 20000 else 2):
                eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            yd = torch.tensor(y, device=dev)
            z, al = eng.potrs_vec(K, dinv, yd)
            bs = ev(lambda: eng.potrs_vec(K, dinv, yd), reps=5)
            print(json.dumps({"op": "potrf_pdl", "n": n, "pdl": pdl, "info": info, "logdet": ld, "ms": round(best, 4),
                              "potrs_vec_ms": round(bs, 4), "alpha_sum": float(al.sum()), "z_sq": float(z @ z)}), flush=True)
        eng.set("pdl", 1)
        del K
if "oz2" in what:
    # modular (CRT) int8 product, first multicast-free form, against the digit-plane kernel and the DMMA kernel
    for (M, N, K) in ((4096, 4096, 2048), (8192, 8192, 2048), (8192, 8192, 512)):
        A = torch.randn(M, K, dtype=torch.float64, device=dev); B = torch.randn(N, K, dtype=torch.float64, device=dev)
        C1 = torch.zeros(M, N, dtype=torch.float64, device=dev); C2 = torch.zeros_like(C1); C3 = torch.zeros_like(C1)
        ms2 = ev(lambda: eng.oz2_gemm(A, B, C1, alpha=-1.0), reps=3)
        ms1 = ev(lambda: eng.gemm_nt_i8(A, B, C2, alpha=-1.0), reps=3)
        ms0 = ev(lambda: eng.gemm_nt(A, B, C3, alpha=-1.0, beta=1.0), reps=3)
        rel = float((C1 / 4 - C3 / 4).norm() / (C3 / 4).norm())
        print(json.dumps({"op": "oz2_gemm", "M": M, "N": N, "K": K, "ms_modular_incl_slice_crt": round(ms2, 3), "ms_digitplanes_incl_slice": round(ms1, 3),
                          "ms_dmma": round(ms0, 3), "tflops_equiv_modular": round(2 * M * N * K / ms2 * 1e-9, 2),
                          "tflops_equiv_digitplanes": round(2 * M * N * K / ms1 * 1e-9, 2), "rel_diff_vs_dmma": rel}), flush=True)
        del A, B, C1, C2, C3

if "sched2" in what:
    # round 2: first-panel width, tail policy (persistent trailing updates on nsm - reserve SMs below `sched_tail` rows), tiles per CTA
    spec = E.battgp_spec()
    DEF = {"sched_t1024": 9000, "sched_t2048": 17000, "sched_t4096": 0, "sched_w0": 0, "sched_w1": 0, "sched_tail": 0, "oz_reserve": 0, "oz_tpc": 2}
    def run(n, xd, K, cfg, reps=2):
        for k, v in DEF.items(): eng.set(k, cfg.get(k, v))
        eng.set("nb", cfg.get("nb", 0)); eng.set("ozaki", 1)
        best = 1e30
        for r in range(reps):
            eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(json.dumps({"op": "potrf_sched2", "n": n, "cfg": cfg, "info": info, "logdet": ld, "ms": round(best, 3)}), flush=True)
    cfgs = [{}, {"sched_w0": 1024}, {"sched_w0": 512}, {"sched_w0": 1024, "sched_w1": 1024},
            {"sched_tail": 16000, "oz_reserve": 8}, {"sched_tail": 16000, "oz_reserve": 16}, {"sched_tail": 16000, "oz_reserve": 32},
            {"sched_tail": 24000, "oz_reserve": 16}, {"sched_tail": 10000, "oz_reserve": 16}, {"sched_tail": 10000, "oz_reserve": 32},
            {"sched_tail": 50000, "oz_reserve": 8}, {"sched_tail": 50000, "oz_reserve": 16},
            {"sched_t2048": 12000}, {"sched_t2048": 22000}, {"sched_t1024": 5000}, {"sched_t1024": 12000},
            {"oz_tpc": 4}, {"oz_tpc": 3}, {"sched_t4096": 30000}, {"sched_w0": 1024, "sched_tail": 16000, "oz_reserve": 16}]
    for n in (40000, 16384):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev)
        K = E.alloc_matrix(n, n, dev)
        for cfg in cfgs: run(n, xd, K, cfg)
        del K
    for k, v in DEF.items(): eng.set(k, v)
    eng.set("nb", 0)
if "potrf2" in what:
    # round 2: default schedule at the sizes of the small-N / mid-N targets, two-stream leaf chain on/off, cuSOLVER (torch) beside it
    spec = E.battgp_spec()
    eng.set("nb", 0); eng.set("lookahead", 1)
    for n in (1024, 2048, 4096, 8192, 16384, 40000):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev)
        K = E.alloc_matrix(n, n, dev)
        row = {"op": "potrf2", "n": n}
        for lc in (1, 0, 1, 0):
            eng.set("leaf_chain", lc)
            best = 1e30
            for r in range(4 if n <= 16384 else 2):
                eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            key = f"ms_leaf_chain{lc}"
            row[key] = round(min(best, row.get(key, 1e30)), 3); row["info"] = info; row[f"logdet{lc}"] = ld
        eng.set("leaf_chain", 1)
        row["tflops_equiv"] = round(n ** 3 / 3 / row["ms_leaf_chain1"] * 1e-9, 2)
        if n <= 16384:
            eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
            Kc = K[:, :n].clone()
            t_copy = ev(lambda: Kc.copy_(K[:, :n]), reps=3)
            t_both = ev(lambda: torch.linalg.cholesky_ex(Kc.copy_(K[:, :n])), reps=3)
            row["cusolver_potrf_ms"] = round(t_both - t_copy, 3)
            del Kc
        print(json.dumps(row), flush=True)
        del K
        torch.cuda.empty_cache()

if "chainwhole" in what:
    spec = E.battgp_spec()
    eng.set("nb", 0); eng.set("lookahead", 1); eng.set("leaf_chain", 1)
    for n in (1536, 2048, 3072, 4096, 6144):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev)
        K = E.alloc_matrix(n, n, dev)
        row = {"op": "chainwhole", "n": n}
        for cw in (0, 8192, 0, 8192):
            eng.set("leaf_chain_max", 8192 if cw else 5120); eng.set("chain_whole_max", cw)
            best = 1e30
            for r in range(4):
                eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            key = "ms_whole_chain" if cw else "ms_panels"
            row[key] = round(min(best, row.get(key, 1e30)), 3); row["info"] = info; row["logdet_" + key] = ld
        print(json.dumps(row), flush=True)
        del K
    eng.set("leaf_chain_max", 5120); eng.set("chain_whole_max", 5120)

if "chaincfg" in what:
    spec = E.battgp_spec()
    eng.set("nb", 0); eng.set("lookahead", 1); eng.set("leaf_chain", 1)
    for n in (1024, 2048, 4096, 8192, 16384):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev)
        K = E.alloc_matrix(n, n, dev)
        row = {"op": "chaincfg", "n": n}
        for cc in (0, 1, 0, 1):
            eng.set("chain_cfg", cc)
            best = 1e30
            for r in range(4):
                eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            key = f"ms_chain_cfg{cc}"
            row[key] = round(min(best, row.get(key, 1e30)), 3); row["info"] = info; row["logdet" + str(cc)] = ld
        print(json.dumps(row), flush=True)
        del K
    eng.set("chain_cfg", 1)

if "chaingemm" in what:
    # which tile form should the side-stream rank-128 updates of the leaf chain use?
    spec = E.battgp_spec()
    eng.set("nb", 0); eng.set("lookahead", 1); eng.set("leaf_chain", 1)
    for n in (2048, 4096, 5000):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev)
        K = E.alloc_matrix(n, n, dev)
        row = {"op": "chaingemm", "n": n}
        for rep in range(2):
            for cfg in (0, 1, 4, 6):
                eng.set("gemm_cfg", cfg)
                best = 1e30
                for r in range(4):
                    eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
                    torch.cuda.synchronize()
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                key = f"ms_gemm_cfg{cfg}"
                row[key] = round(min(best, row.get(key, 1e30)), 3)
        print(json.dumps(row), flush=True)
        del K
    eng.set("gemm_cfg", 0)

if "chainsplit" in what:
    spec = E.battgp_spec()
    eng.set("nb", 0); eng.set("lookahead", 1); eng.set("leaf_chain", 1); eng.set("chain_cfg", 1)
    for n in (1000, 1024, 2048, 4096, 8192, 16384, 40000):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev)
        K = E.alloc_matrix(n, n, dev)
        row = {"op": "chainsplit", "n": n}
        for cc in (0, 1, 0, 1):
            eng.set("chain_split", cc)
            best = 1e30
            for r in range(4 if n <= 16384 else 2):
                eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            key = f"ms_split{cc}"
            row[key] = round(min(best, row.get(key, 1e30)), 3); row["info"] = info; row["logdet" + str(cc)] = ld
        print(json.dumps(row), flush=True)
        del K
    eng.set("chain_split", 0)

if "pdlrows" in what:
    spec = E.battgp_spec()
    eng.set("nb", 0); eng.set("lookahead", 1)
    for n in (8192, 16384, 24576, 40000):
        x, y = synth_field_data(n, 0)
        xd = torch.tensor(x, device=dev)
        K = E.alloc_matrix(n, n, dev)
        row = {"op": "pdlrows", "n": n}
        for rep in range(2):
            for pr in (12000, 1000000):
                eng.set("pdl_chain_rows", pr)
                best = 1e30
                for r in range(3 if n <= 16384 else 2):
                    eng.cov_build(spec, xd, noise=2.33e-6, symmetric=True, out=K)
                    torch.cuda.synchronize()
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(); info, ld, dinv = eng.potrf(K); e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                key = f"ms_pdl_rows_{pr}"
                row[key] = round(min(best, row.get(key, 1e30)), 3); row["info"] = info
        print(json.dumps(row), flush=True)
        del K
    eng.set("pdl_chain_rows", 12000)
