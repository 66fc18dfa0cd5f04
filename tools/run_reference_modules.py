"""Run the reference's OWN, UNMODIFIED modules on the GPU through the import shim and check what they return.

    python tools/run_reference_modules.py [--out profiles/reference_modules_r02.jsonl]

What runs (all imported from the reference tree, nothing re-typed):
  (i)   src/batt_models/battcellgp_full.py  BatteryCellGP_Full(x, y, device="cuda").predict(Xq) / .predict_r0_op(op, t)
        on the four real-field-data sets of tests/golden/real_field_data.npz, against the committed CPU results
        (mean rtol 1e-6, variance rtol 1e-4, DataFrame columns t / r0_acausal_<tag> / r0var_acausal_<tag>);
  (ii)  src/gp/training.py  train_exact_gp_adam / train_exact_gp_lbfgs / train_exact_gp_botorch on a BatteryCellGP, against a
        CPU replica of the same optimiser fed by the oracle's LML and analytic gradient (losses per iteration rtol 1e-6);
        BatteryCellGP_Full.train_hyperparameters (battcellgp_full.py:127-166) end to end;
  (iii) src/gp/standard_models.py  ScaledRBFModel.predict (fp32 in/out) against the oracle;
  (iv)  the reference's own unit tests tests/gp/test_standard_models.py, test_recursive_gp.py, test_spatiotemporal_gp.py,
        test_wiener_temporal_kernel.py through unittest.

The reference tree is found at $BATTGP_REFERENCE, /root/reference (build container) or unpacked from
oracle/_ref/reference_src.tar.gz (written by tools/stage_reference.py; git-ignored, travels with the gpurun snapshot --
reference sources never enter this repository's history).  The oracle is used as the checker only.
"""
from __future__ import annotations

import argparse
import copy
import io
import json
import math
import os
import sys
import tarfile
import tempfile
import unittest
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TARBALL = os.path.join(ROOT, "oracle", "_ref", "reference_src.tar.gz")


def find_reference():
    """Path of a reference tree (directory holding src/), or None."""
    for cand in (os.environ.get("BATTGP_REFERENCE"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "src", "batt_models")):
            return cand
    if os.path.exists(TARBALL):
        dst = os.path.join(tempfile.gettempdir(), "battgp_reference_%d" % os.getuid())
        if not os.path.isdir(os.path.join(dst, "src", "batt_models")):
            # several ranks may get here at once: unpack privately, then publish with one atomic rename
            tmp = tempfile.mkdtemp(prefix="battgp_reference_unpack_")
            with tarfile.open(TARBALL) as tf:
                tf.extractall(tmp, filter="data")
            try:
                os.rename(tmp, dst)
            except OSError:                      # another process published first
                import shutil
                shutil.rmtree(tmp, ignore_errors=True)
        return dst
    return None


def install(ref):
    import battgp_b200.shim as shim
    shim.install(force=True)
    if ref not in sys.path:
        sys.path.insert(0, ref)


# ------------------------------------------------------------------------------------------------------------ (i)
def check_predict(dev, emit):
    import numpy as np
    import torch
    from src.batt_models.battcellgp_full import BatteryCellGP_Full
    from src.operating_point import Op
    g = np.load(os.path.join(ROOT, "tests", "golden", "real_field_data.npz"))
    ok = True
    for key, cellnr in (("b14_c1", 1), ("b14_cpack", -1), ("b3_c5", 5), ("b14_c3", 3)):
        x, y, xq = g[key + "_x"], g[key + "_y"], g[key + "_xq"]
        cell = BatteryCellGP_Full(x, y, cellnr, device=dev)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mean, var = cell.predict(xq)
            op = Op(float(xq[0, 1]), float(xq[0, 2]), float(xq[0, 3]))
            df = cell.predict_r0_op(op, xq[:, 0])
        tag = "pack" if cellnr == -1 else f"c{cellnr}"
        cols_ok = list(df.columns) == ["t", f"r0_acausal_{tag}", f"r0var_acausal_{tag}"]
        r = {"check": "BatteryCellGP_Full.predict/predict_r0_op", "set": key, "n": int(x.shape[0]), "device": str(dev),
             "mean_max_rel": float(np.max(np.abs(mean - g[key + "_mean"]) / np.abs(g[key + "_mean"]))),
             "var_max_rel": float(np.max(np.abs(var - g[key + "_var"]) / np.abs(g[key + "_var"]))),
             "dataframe_columns": list(df.columns), "columns_ok": cols_ok,
             "df_equals_predict": bool(np.array_equal(df.iloc[:, 1].to_numpy(), mean) and np.array_equal(df.iloc[:, 2].to_numpy(), var)),
             "returns_numpy_float64": bool(isinstance(mean, np.ndarray) and mean.dtype == np.float64 and var.dtype == np.float64),
             "training_data_roundtrip": bool(np.array_equal(cell.get_training_data()[0], x)),
             "tolerance": {"mean": 1e-6, "var": 1e-4}}
        r["ok"] = bool(r["mean_max_rel"] < 1e-6 and r["var_max_rel"] < 1e-4 and cols_ok and r["df_equals_predict"]
                       and r["returns_numpy_float64"] and r["training_data_roundtrip"])
        ok &= r["ok"]
        emit(r)
        del cell
    return ok


# ------------------------------------------------------------------------------------------------------------ (ii)
def _replica(model, x, y):
    """CPU copies of the raw parameters / constraints of a BatteryCellGP and a closure that evaluates -LML/N and its
    gradient with respect to the RAW parameters through the oracle (analytic dLML/dtheta) and the constraint transforms."""
    import numpy as np
    import torch
    from oracle import gp_oracle as orc
    k0, k1 = model.covar_module.kernels[0], model.covar_module.kernels[1]
    owners = [(model.likelihood.noise_covar, "raw_noise"), (k0, "raw_outputscale"), (k1, "raw_outputscale"),
              (k1.base_kernel, "raw_lengthscale")]
    raws, cons = [], []
    for mod, name in owners:
        raws.append(getattr(mod, name).detach().cpu().double().clone().requires_grad_(True))
        cons.append(copy.deepcopy(getattr(mod, name + "_constraint")).cpu())
    n = x.shape[0]

    def theta():
        return [c.transform(r) for c, r in zip(cons, raws)]

    def evaluate(backward=True):
        th = theta()
        noise, s_w, s_r = (float(t.detach().reshape(-1)[0]) for t in th[:3])
        ls = [float(v) for v in th[3].detach().reshape(-1)]
        spec = orc.battgp_spec(s_w, s_r, ls)
        res = orc.lml_grad(spec, x, y, noise)
        if backward:
            g = [res["noise"], res["terms"][0]["outputscale"], res["terms"][1]["outputscale"]]
            gl = torch.tensor(res["terms"][1]["lengthscale"], dtype=torch.float64).reshape(th[3].shape)
            sur = sum((-gi / n) * t.sum() for gi, t in zip(g, th[:3])) + ((-gl / n) * th[3]).sum()
            sur.backward()
        return -res["lml"] / n

    return raws, evaluate


def check_training(dev, emit):
    import numpy as np
    import torch
    from scipy.optimize import minimize
    from src.batt_models.battcellgp_full import BatteryCellGP_Full
    from src.gp import training
    g = np.load(os.path.join(ROOT, "tests", "golden", "real_field_data.npz"))
    x, y = g["b3_c5_x"][::2][:400].copy(), g["b3_c5_y"][::2][:400].copy()
    n = x.shape[0]
    ok = True

    def fresh():
        cell = BatteryCellGP_Full(x, y, 5, device=dev)
        return cell, cell.model.train_inputs[0], cell.model.train_targets

    # ---- Adam (training.py:11-67), 5 iterations, no early stop
    cell, tx, ty = fresh()
    raws, evaluate = _replica(cell.model, x, y)
    iters, lr = 5, 0.1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        losses = training.train_exact_gp_adam(cell.model, tx, ty, max_iter=iters, rel_ftol=0.0, loss_scale=n, lr=lr, messages=False)
    opt = torch.optim.Adam(raws, lr=lr)
    ref = []
    for _ in range(iters):
        opt.zero_grad()
        ref.append(evaluate() * n)
        opt.step()
    ref_final = evaluate(backward=False) * n
    r = {"check": "training.train_exact_gp_adam", "n": n, "iters": iters, "lr": lr, "losses": [float(v) for v in losses],
         "replica_losses": ref + [ref_final],
         "max_rel_diff_per_iteration": float(np.max(np.abs(np.asarray(losses[:iters]) - np.asarray(ref)) / np.abs(ref))),
         "returned_shape_ok": bool(isinstance(losses, np.ndarray) and losses.shape == (iters + 1,)),
         "final_loss_rel_diff_vs_loss_after_last_step": abs(losses[-1] - ref_final) / abs(ref_final),
         "final_loss_rel_diff_vs_last_iteration": abs(losses[-1] - ref[-1]) / abs(ref[-1]),
         "model_back_in_eval_mode": bool(not cell.model.training), "tolerance": 1e-6}
    r["ok"] = bool(r["max_rel_diff_per_iteration"] < 1e-6 and r["returned_shape_ok"] and r["model_back_in_eval_mode"]
                   and min(r["final_loss_rel_diff_vs_loss_after_last_step"], r["final_loss_rel_diff_vs_last_iteration"]) < 1e-6)
    ok &= r["ok"]; emit(r)
    # the rel_ftol stopping rule (training.py:47-53): a huge tolerance stops at i = 1 and truncates the array to i + 1 entries
    cell, tx, ty = fresh()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        l2 = training.train_exact_gp_adam(cell.model, tx, ty, max_iter=10, rel_ftol=1e9, loss_scale=n, lr=lr, messages=False)
    r = {"check": "training.train_exact_gp_adam rel_ftol stop", "returned_length": int(len(l2)), "expected_length": 2}
    r["ok"] = bool(len(l2) == 2)
    ok &= r["ok"]; emit(r)

    # ---- torch L-BFGS with strong Wolfe (training.py:108-171), 3 iterations
    cell, tx, ty = fresh()
    raws, evaluate = _replica(cell.model, x, y)
    iters = 3
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        losses = training.train_exact_gp_lbfgs(cell.model, tx, ty, max_iter=iters, rel_ftol=0.0, loss_scale=n, lr=1.0, messages=False)
    opt = torch.optim.LBFGS(raws, line_search_fn="strong_wolfe", lr=1.0)

    def closure():
        opt.zero_grad()
        return torch.tensor(evaluate())
    ref = []
    for _ in range(iters):
        ref.append(evaluate(backward=False) * n)
        opt.step(closure)
    r = {"check": "training.train_exact_gp_lbfgs", "n": n, "iters": iters, "losses": [float(v) for v in losses], "replica_losses": ref,
         "max_rel_diff_per_iteration": float(np.max(np.abs(np.asarray(losses[:iters]) - np.asarray(ref)) / np.abs(ref))),
         "loss_decreased": bool(losses[iters - 1] < losses[0]), "tolerance": 1e-5}
    r["ok"] = bool(r["max_rel_diff_per_iteration"] < 1e-5 and r["loss_decreased"])
    ok &= r["ok"]; emit(r)

    # ---- botorch stand-in (training.py:70-105): scipy L-BFGS-B to convergence
    cell, tx, ty = fresh()
    raws, evaluate = _replica(cell.model, x, y)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        buf = io.StringIO()
        so, sys.stdout = sys.stdout, buf
        try:
            final = training.train_exact_gp_botorch(cell.model, tx, ty)
        finally:
            sys.stdout = so
    sizes = [t.numel() for t in raws]

    def fun(v):
        o = 0
        with torch.no_grad():
            for t, sz in zip(raws, sizes):
                t.copy_(torch.as_tensor(v[o:o + sz]).reshape(t.shape)); o += sz
        for t in raws:
            t.grad = None
        try:
            val = evaluate()
        except Exception:
            return 1e10, np.zeros_like(v)
        return val, np.concatenate([t.grad.reshape(-1).numpy() for t in raws])
    x0 = np.concatenate([t.detach().reshape(-1).numpy() for t in raws])
    l0 = fun(x0)[0]
    res = minimize(fun, x0, jac=True, method="L-BFGS-B", options={"maxiter": 10000, "ftol": 1e-15, "gtol": 1e-15, "maxfun": 10000, "maxls": 10000})
    # training.py:100-105 returns -mll of the EVAL-mode output (the posterior at the training inputs), not the training loss:
    # the fitted hyper-parameters are what is compared -- the oracle's -LML/N at them against the replica's optimum
    from oracle import gp_oracle as orc
    m = cell.model
    spec = orc.battgp_spec(float(m.outputscale_wiener), float(m.outputscale_rbf), [float(v) for v in m.lengthscale_rbf.detach().cpu().reshape(-1)])
    trained = -orc.fit(spec, x, y, float(m.noise_variance)).lml / n
    r = {"check": "training.train_exact_gp_botorch", "n": n, "start_loss": l0, "loss_at_fitted_hyperparameters": float(trained),
         "replica_final_loss": float(res.fun), "rel_diff": abs(trained - res.fun) / abs(res.fun),
         "returned_value_minus_mll_of_eval_output": float(final), "returned_float": isinstance(final, float), "tolerance": 1e-4}
    r["ok"] = bool(r["rel_diff"] < 1e-4 and trained < l0 and r["returned_float"] and math.isfinite(final))
    ok &= r["ok"]; emit(r)

    # ---- BatteryCellGP_Full.train_hyperparameters (battcellgp_full.py:127-166) with the config's trainer (Adam, lr 1)
    cell, _, _ = fresh()
    cell.params["max_iter"] = 4
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        l3 = cell.train_hyperparameters(messages=False)
    p = cell.get_parameters()
    r = {"check": "BatteryCellGP_Full.train_hyperparameters", "losses": [float(v) for v in np.asarray(l3).reshape(-1)],
         "params_after": {k: (list(map(float, p[k])) if isinstance(p[k], tuple) else float(p[k]))
                          for k in ("noise_variance", "outputscale_wiener", "outputscale_rbf", "lengthscale_rbf")},
         "marginallikelihood": float(cell.marginallikelihood)}
    r["ok"] = bool(np.all(np.isfinite(np.asarray(l3))) and len(p["lengthscale_rbf"]) == 3 and math.isfinite(cell.marginallikelihood))
    ok &= r["ok"]; emit(r)
    return ok


# ------------------------------------------------------------------------------------------------------------ (iii)
def check_scaled_rbf(dev, emit):
    import numpy as np
    import torch
    from oracle import gp_oracle as orc
    from src.gp.standard_models import ScaledRBFModel
    rng = np.random.default_rng(3)
    x = rng.normal(size=(200, 3)); y = np.sin(x[:, 0]) + 0.1 * rng.normal(size=200); xq = rng.normal(size=(40, 3))
    gp = ScaledRBFModel(torch.tensor(x, device=dev), torch.tensor(y, device=dev), noise_variance=0.05, outputscale=1.3, lengthscale=0.9)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m, v = gp.predict(torch.tensor(xq, device=dev))
        m2, c2 = gp.predict(torch.tensor(xq, device=dev), full_cov=True)
    spec = orc.scaled_rbf_spec(3, 1.3, 0.9)
    x32, y32, xq32 = x.astype(np.float32).astype(np.float64), y.astype(np.float32).astype(np.float64), xq.astype(np.float32).astype(np.float64)
    f = orc.fit(spec, x32, y32, float(np.float32(0.05)))
    mr, cr = orc.predict(spec, x32, f, xq32, full_cov=True)
    m, v, c2 = (np.asarray(t.detach().cpu() if hasattr(t, "detach") else t, dtype=np.float64) for t in (m, v, c2))
    r = {"check": "ScaledRBFModel.predict (fp32 model, standard_models.py:8-55)", "mean_max_abs": float(np.max(np.abs(m - mr))),
         "var_max_abs": float(np.max(np.abs(v - np.diag(cr)))), "full_cov_rel_fro": float(np.linalg.norm(c2 - cr) / np.linalg.norm(cr)),
         "tolerance": {"mean_abs": 1e-4, "var_abs": 1e-4, "cov_rel_fro": 1e-4}}
    r["ok"] = bool(r["mean_max_abs"] < 1e-4 and r["var_max_abs"] < 1e-4 and r["full_cov_rel_fro"] < 1e-4)
    emit(r)
    ok = r["ok"]
    # SparseScaledRBFModel (standard_models.py:58-107, SGPR through InducingPointKernel): subset-of-regressors prediction
    from src.gp.standard_models import SparseScaledRBFModel
    u = x[rng.choice(200, 25, replace=False)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sgp = SparseScaledRBFModel(torch.tensor(x), torch.tensor(y), torch.tensor(u), 0.05, 1.3, 0.9)
        ms, vs = sgp.predict(xq)
    u32 = u.astype(np.float32).astype(np.float64)
    nz = float(np.float32(0.05))
    Kuu = orc.cov(spec, u32, u32); Kfu = orc.cov(spec, x32, u32); Ksu = orc.cov(spec, xq32, u32)
    Qff = Kfu @ np.linalg.solve(Kuu, Kfu.T); Qsf = Ksu @ np.linalg.solve(Kuu, Kfu.T); Qss = Ksu @ np.linalg.solve(Kuu, Ksu.T)
    G = Qff + nz * np.eye(200)
    mref = Qsf @ np.linalg.solve(G, y32)
    vref = np.diag(Qss - Qsf @ np.linalg.solve(G, Qsf.T))
    r = {"check": "SparseScaledRBFModel.predict (SGPR, standard_models.py:58-107)", "mean_max_abs": float(np.max(np.abs(ms - mref))),
         "var_max_abs": float(np.max(np.abs(vs - vref))), "tolerance": {"mean_abs": 2e-4, "var_abs": 2e-4}}
    r["ok"] = bool(r["mean_max_abs"] < 2e-4 and r["var_max_abs"] < 2e-4)
    emit(r)
    return bool(ok and r["ok"])


# ------------------------------------------------------------------------------------------------------------ (iv)
def run_reference_unit_tests(ref, emit):
    ok = True
    for mod in ("tests.gp.test_standard_models", "tests.gp.test_recursive_gp", "tests.gp.test_spatiotemporal_gp",
                "tests.gp.test_wiener_temporal_kernel"):
        if not os.path.exists(os.path.join(ref, *mod.split(".")) + ".py"):
            emit({"check": "reference unit tests", "module": mod, "skipped": "file not staged"})
            continue
        for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
            del sys.modules[k]                    # this repository also has a top-level `tests` package
        old = sys.path[:]
        sys.path.insert(0, ref)
        try:
            # the reference's tests draw UNSEEDED random inputs, and tests/gp/test_spatiotemporal_gp.py::test_compare_stgp_rgp
            # pushes them through pinv(K_b, rcond=1e-8) of a rank-deficient matrix (recursive_gp.py:63): it fails for ~7 % of
            # the draws even in plain CPU torch (60 draws against the stand-in gpytorch of tests/golden/make_golden.py: 4
            # failures).  A fixed seed makes the run reproducible; seed 0 passes on that CPU stand-in.
            import numpy as _np
            import torch as _torch
            _np.random.seed(0)
            _torch.manual_seed(0)
            suite = unittest.defaultTestLoader.loadTestsFromName(mod)
            buf = io.StringIO()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = unittest.TextTestRunner(stream=buf, verbosity=0).run(suite)
            r = {"check": "reference unit tests (unmodified, through the shim)", "module": mod, "rng_seed": 0, "run": res.testsRun,
                 "failures": [str(t[0]) for t in res.failures], "errors": [str(t[0]) for t in res.errors],
                 "error_text": [t[1][-600:] for t in (res.failures + res.errors)][:4], "skipped": [str(t[0]) for t in res.skipped]}
            r["ok"] = bool(res.wasSuccessful() and res.testsRun > 0)
        except Exception as e:  # import error etc.
            r = {"check": "reference unit tests (unmodified, through the shim)", "module": mod, "ok": False, "exception": repr(e)}
        finally:
            sys.path[:] = old
        ok &= r["ok"]
        emit(r)
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip", default="", help="comma list of: predict,training,rbf,unittests")
    args = ap.parse_args()
    ref = find_reference()
    if ref is None:
        print(json.dumps({"error": "no reference tree: set BATTGP_REFERENCE or run tools/stage_reference.py in the build container"}))
        return 2
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; battgp_b200 has no CPU fallback"}))
        return 2
    torch.set_default_dtype(torch.float64)        # what gp_runner.py's worker does (gp_runner.py:158-159)
    install(ref)
    dev = torch.device("cuda", 0)
    out = open(args.out, "w") if args.out else None
    results = []

    def emit(r):
        results.append(r)
        line = json.dumps(r)
        print(line, flush=True)
        if out:
            out.write(line + "\n"); out.flush()
    skip = set(args.skip.split(","))
    emit({"reference_root": ref, "gpytorch": sys.modules["gpytorch"].__version__, "device": torch.cuda.get_device_name(0)})
    ok = True
    if "predict" not in skip:
        ok &= check_predict(dev, emit)
    if "training" not in skip:
        ok &= check_training(dev, emit)
    if "rbf" not in skip:
        ok &= check_scaled_rbf(dev, emit)
    if "unittests" not in skip:
        ok &= run_reference_unit_tests(ref, emit)
    from battgp_b200 import engine as E
    emit({"summary": "reference modules on the GPU through the shim", "all_ok": bool(ok), "checks": len(results) - 1,
          "failed": [r.get("check") or r.get("module") for r in results if r.get("ok") is False],
          "kernel_launches": E.get_engine(dev).launches})
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
