"""CPU reference arm for BattGP's ``full_gp`` path in plain torch fp64 -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/`` and ``bench.py``'s ``--impl reference`` / ``cpu_baseline`` legs may import this module; nothing under
``battgp_b200/`` does.

This is the path BASELINE.md section 3 prescribes as "the reference CPU path": what GPyTorch executes for
``BatteryCellGP_Full(x, y).predict(X*)`` (/root/reference/src/batt_models/battcellgp_full.py:168-195) under
``gpytorch.settings.max_cholesky_size(N+1)``, written with the torch ops GPyTorch itself calls
[GPyTorch-recall, SURVEY.md Appendix A.2/C -- GPyTorch is not installable here]:

* RBF-ARD ........ ``x/l``, subtract the column mean, one GEMM on ``[-2a, |a|^2, 1] x [b, 1, |b|^2]^T``, zero the diagonal of
                   the train block, ``clamp_min_(0)``, ``div_(-2)``, ``exp_()``      (cell_gp.py:33; gpytorch Kernel.covar_dist)
* Wiener ......... ``min^3/3 + |t-t'| min^2/2``                                     (/root/reference/src/gp/wiener_kernel.py:32;
                   vectorised -- the reference fills ``min`` with a Python loop over columns, :21-22)
* scale, sum, noise ``s_w K_W + s_r K_R + sigma_n^2 I``                             (cell_gp.py:27,34-36)
* factorisation .. ``torch.linalg.cholesky_ex``                                     (psd_safe_cholesky)
* alpha .......... ``torch.cholesky_solve``                                         (prediction-strategy mean cache)
* variance ....... ``torch.linalg.solve_triangular`` with the M = 300 right-hand sides, ``k** - colsum(V^2)``, clamp 1e-10
* LML ............ ``-0.5 y.alpha - sum log L_ii - N/2 log 2 pi``                   (/root/reference/src/gp/training.py:27-43)

The N x N build runs in row blocks (same torch ops per block) so that no N^2-sized temporaries pile up: N = 40 000 needs
K (12.8 GB) + L (12.8 GB) + ~2 GB.  Results are checked against ``oracle/gp_oracle.py`` and the golden vectors in
tests/test_oracle.py::test_torch_reference_arm_matches_oracle.
"""
from __future__ import annotations

import math
import time

import torch

MIN_VARIANCE_F64 = 1e-10


def _rbf_block(a: torch.Tensor, b: torch.Tensor, ls: torch.Tensor, adj: torch.Tensor, diag_offset=None) -> torch.Tensor:
    """exp(-0.5 sqdist) for rows ``a`` against ``b`` by GPyTorch's centred quadratic expansion."""
    a_ = a / ls - adj
    b_ = b / ls - adj
    an = a_.pow(2).sum(1, keepdim=True)
    bn = b_.pow(2).sum(1, keepdim=True)
    lhs = torch.cat([-2.0 * a_, an, torch.ones_like(an)], dim=1)
    rhs = torch.cat([b_, torch.ones_like(bn), bn], dim=1)
    res = lhs @ rhs.T
    if diag_offset is not None:                      # x1 is x2: GPyTorch zero-fills the diagonal
        res.diagonal(offset=diag_offset).zero_()
    return res.clamp_min_(0.0).div_(-2.0).exp_()


def _wiener_block(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    t1 = t1.reshape(-1, 1)
    t2 = t2.reshape(1, -1)
    m = torch.minimum(t1, t2)
    d = (t1 - t2).abs_()
    return m.pow(3).div_(3.0).add_(d.mul_(m.pow(2)).div_(2.0))


def cov(x1: torch.Tensor, x2: torch.Tensor, s_w: float, s_r: float, ls, *, noise: float = 0.0, same: bool = False,
        out: torch.Tensor | None = None, block: int = 4096) -> torch.Tensor:
    """``s_w Wiener(t) + s_r RBF-ARD(I, SOC, T)`` (+ noise on the diagonal when ``same``), built block-row-wise."""
    n1, n2 = x1.shape[0], x2.shape[0]
    ls_t = torch.as_tensor(ls, dtype=torch.float64).reshape(1, -1)
    adj = (x1[:, 1:4] / ls_t).mean(0, keepdim=True)
    if out is None:
        out = torch.empty((n1, n2), dtype=torch.float64)
    for r0 in range(0, n1, block):
        r1 = min(n1, r0 + block)
        kb = _rbf_block(x1[r0:r1, 1:4], x2[:, 1:4], ls_t, adj, diag_offset=r0 if same else None).mul_(s_r)
        kb.add_(_wiener_block(x1[r0:r1, 0], x2[:, 0]).mul_(s_w))
        if same:
            kb.diagonal(offset=r0).add_(noise)
        out[r0:r1].copy_(kb)
    return out


def fit_predict(x, y, xq, *, noise: float = 2.33e-6, s_w: float = 4.23e-13, s_r: float = 0.0099,
                ls=(12.11, 33.75, 45.14), clamp: bool = True) -> dict:
    """One fit + predict pass; returns mean, var, lml, alpha and the seconds spent per phase (time.perf_counter)."""
    x = torch.as_tensor(x, dtype=torch.float64)
    y = torch.as_tensor(y, dtype=torch.float64).reshape(-1)
    xq = torch.as_tensor(xq, dtype=torch.float64)
    n = x.shape[0]
    ph = {}
    t0 = time.perf_counter()
    K = cov(x, x, s_w, s_r, ls, noise=noise, same=True)
    t1 = time.perf_counter(); ph["build"] = t1 - t0
    L, info = torch.linalg.cholesky_ex(K)
    del K
    if int(info) != 0:
        raise RuntimeError(f"cholesky_ex info={int(info)}")
    t2 = time.perf_counter(); ph["cholesky"] = t2 - t1
    alpha = torch.cholesky_solve(y[:, None], L)[:, 0]
    t3 = time.perf_counter(); ph["alpha"] = t3 - t2
    lml = -0.5 * float(y @ alpha) - float(L.diagonal().log().sum()) - 0.5 * n * math.log(2.0 * math.pi)
    t4 = time.perf_counter(); ph["lml"] = t4 - t3
    Kq = cov(xq, x, s_w, s_r, ls)
    mean = Kq @ alpha
    V = torch.linalg.solve_triangular(L, Kq.T, upper=False)
    kdiag = s_w * xq[:, 0].pow(3) / 3.0 + s_r
    var = kdiag - V.pow(2).sum(0)
    if clamp:
        var = var.clamp_min(MIN_VARIANCE_F64)
    t5 = time.perf_counter(); ph["predict"] = t5 - t4
    ph["total"] = t5 - t0
    return {"mean": mean.numpy(), "var": var.numpy(), "lml": lml, "alpha": alpha.numpy(), "seconds": ph}
