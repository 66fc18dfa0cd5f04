"""botorch.fit.fit_gpytorch_mll: scipy L-BFGS-B over the raw hyper-parameters (training.py:82-95)."""
from __future__ import annotations

import warnings

import numpy as np
import torch
from scipy.optimize import minimize

from ..engine import NanError, NotPSDError, NumericalWarning


def fit_gpytorch_mll(mll, max_retries: int = 5, optimizer_kwargs=None, **kwargs):
    """The reference calls this with max_retries=10000, maxls=10000, ftol=gtol=1e-15 (training.py:82-95): a line-search probe
    can land where K is not positive definite even after the jitter retries.  Like botorch's ``_fit_fallback`` that must not
    abort the training: a failed evaluation is reported to L-BFGS-B as a large finite loss (the line search backs off), and
    if the optimiser itself gives up on such a point the best parameters seen so far are restored and the fit is retried
    (up to ``max_retries`` times, with a warning)."""
    model = mll.model
    model.train(); model.likelihood.train()
    params = [p for p in model.parameters() if p.requires_grad]
    shapes = [p.shape for p in params]
    sizes = [p.numel() for p in params]
    options = dict((optimizer_kwargs or {}).get("options", {}))
    options.pop("eps", None)                 # finite-difference step: unused, the gradient is analytic
    best = {"loss": np.inf, "x": None, "failed": 0}

    def set_params(xv):
        o = 0
        with torch.no_grad():
            for p, sh, sz in zip(params, shapes, sizes):
                p.copy_(torch.as_tensor(xv[o:o + sz], dtype=p.dtype, device=p.device).reshape(sh))
                o += sz

    def fun(xv):
        set_params(xv)
        for p in params:
            p.grad = None
        try:
            loss = -mll(model(*model.train_inputs), model.train_targets)
            loss.backward()
        except (NotPSDError, NanError) as e:
            best["failed"] += 1
            warnings.warn(f"fit_gpytorch_mll: objective not evaluable at a line-search point ({e}); backing off", NumericalWarning)
            big = 1e10 if not np.isfinite(best["loss"]) else abs(best["loss"]) * 10.0 + 1e3
            return big, np.zeros_like(xv)
        g = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu().double().reshape(-1).numpy()
                            for p in params])
        lv = float(loss.detach().cpu())
        if not np.isfinite(lv) or not np.all(np.isfinite(g)):
            best["failed"] += 1
            return (1e10 if not np.isfinite(best["loss"]) else abs(best["loss"]) * 10.0 + 1e3), np.zeros_like(xv)
        if lv < best["loss"]:
            best["loss"], best["x"] = lv, np.array(xv, copy=True)
        return lv, g

    x0 = np.concatenate([p.detach().cpu().double().reshape(-1).numpy() for p in params])
    res = None
    for attempt in range(max(1, int(max_retries))):
        failed_before = best["failed"]
        res = minimize(fun, x0, jac=True, method="L-BFGS-B", options=options)
        ended_on_failure = best["x"] is not None and (not np.isfinite(res.fun) or res.fun > best["loss"])
        if best["failed"] == failed_before or not ended_on_failure:
            break
        warnings.warn("fit_gpytorch_mll: optimiser stopped at a non-evaluable point; restarting from the best parameters", NumericalWarning)
        x0 = best["x"]
    xfin = res.x if (best["x"] is None or (np.isfinite(res.fun) and res.fun <= best["loss"])) else best["x"]
    set_params(xfin)
    model.eval(); model.likelihood.eval()
    return mll
