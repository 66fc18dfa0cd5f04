"""botorch.fit.fit_gpytorch_mll: scipy L-BFGS-B over the raw hyper-parameters (training.py:82-95)."""
from __future__ import annotations

import numpy as np
import torch
from scipy.optimize import minimize


def fit_gpytorch_mll(mll, max_retries: int = 5, optimizer_kwargs=None, **kwargs):
    model = mll.model
    model.train(); model.likelihood.train()
    params = [p for p in model.parameters() if p.requires_grad]
    shapes = [p.shape for p in params]
    sizes = [p.numel() for p in params]
    options = dict((optimizer_kwargs or {}).get("options", {}))
    options.pop("eps", None)                 # finite-difference step: unused, the gradient is analytic

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
        loss = -mll(model(*model.train_inputs), model.train_targets)
        loss.backward()
        g = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu().double().reshape(-1).numpy()
                            for p in params])
        return float(loss.detach().cpu()), g

    x0 = np.concatenate([p.detach().cpu().double().reshape(-1).numpy() for p in params])
    res = minimize(fun, x0, jac=True, method="L-BFGS-B", options=options)
    set_params(res.x)
    model.eval(); model.likelihood.eval()
    return mll
