"""Generate tests/golden/*.npz by running the REFERENCE's own code (imported from /root/reference) in the build
container.  Committed together with its outputs; not run on the GPU box (the reference tree does not travel).

GPyTorch is not installable here, so the reference modules are imported against a minimal stand-in
``gpytorch`` module defined below.  The stand-in only supplies what these reference files touch -- a ``Kernel`` base
class (``covar_dist`` = Euclidean distance, ``active_dims`` slicing in ``__call__``), RBF / Scale / additive kernels
written from GPyTorch's published formulas, ``ZeroMean`` and the ``settings`` context managers.  What the golden
vectors pin is therefore the reference's *algorithms*:

* ``wiener_*``   : /root/reference/src/gp/wiener_kernel.py:10-32 executed verbatim (``WienerKernel.forward``).
* ``rgp_*``      : /root/reference/src/gp/recursive_gp.py (Huber's recursive GP) with base points == training inputs,
                   which the reference's own test asserts equals the exact GP (tests/gp/test_recursive_gp.py:195-232).
* ``stgp_*``     : /root/reference/src/gp/spatiotemporal_gp.py (Kalman-form spatio-temporal GP, Wiener temporal kernel
                   from wiener_kernel_temporal.py) which the reference asserts equals the exact Wiener+RBF-ARD GP
                   (tests/gp/test_spatiotemporal_gp.py:218-282) -- an algorithm that shares no code with a Cholesky.

Usage:  python tests/golden/make_golden.py
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


# ----------------------------------------------------------------------------------------------- gpytorch stand-in
def _install_stub():
    g = types.ModuleType("gpytorch")

    class Kernel(torch.nn.Module):
        is_stationary = True
        has_lengthscale = False

        def __init__(self, active_dims=None, ard_num_dims=None, **kw):
            super().__init__()
            self.active_dims = None if active_dims is None else torch.tensor(list(active_dims), dtype=torch.long)
            self.ard_num_dims = ard_num_dims

        @property
        def device(self):
            return torch.device("cpu")

        def covar_dist(self, x1, x2, diag=False, square_dist=False, **params):
            if diag:
                d = (x1 - x2).pow(2).sum(-1)
                return d if square_dist else d.sqrt()
            d = torch.cdist(x1, x2)
            return d.pow(2) if square_dist else d

        def __call__(self, x1, x2=None, diag=False, **params):
            x2 = x1 if x2 is None else x2
            if self.active_dims is not None:
                x1 = x1.index_select(-1, self.active_dims)
                x2 = x2.index_select(-1, self.active_dims)
            res = self.forward(x1, x2, diag=diag, **params)
            if diag and res.dim() == 2:            # GPyTorch's diag fallback
                res = res.diagonal()
            return res

        def __add__(self, other):
            return AdditiveKernel(self, other)

        def __getitem__(self, idx):
            return self

    class AdditiveKernel(Kernel):
        def __init__(self, *kernels):
            super().__init__()
            self.kernels = torch.nn.ModuleList(kernels)

        def forward(self, x1, x2, diag=False, **params):
            out = 0
            for k in self.kernels:
                out = out + k(x1, x2, diag=diag, **params)
            return out

    class RBFKernel(Kernel):
        has_lengthscale = True

        def __init__(self, ard_num_dims=None, active_dims=None, **kw):
            super().__init__(active_dims=active_dims, ard_num_dims=ard_num_dims)
            self.lengthscale = torch.ones(1, ard_num_dims or 1, dtype=torch.float64)

        def forward(self, x1, x2, diag=False, **params):
            ls = torch.as_tensor(self.lengthscale, dtype=x1.dtype).reshape(1, -1)
            a, b = x1 / ls, x2 / ls
            if diag:
                return torch.exp(-0.5 * (a - b).pow(2).sum(-1))
            return torch.exp(-0.5 * torch.cdist(a, b).pow(2))

    class ScaleKernel(Kernel):
        def __init__(self, base_kernel, **kw):
            super().__init__(active_dims=None)
            self.base_kernel = base_kernel
            self.active_dims = base_kernel.active_dims      # GPyTorch copies active_dims from the base kernel
            self.outputscale = torch.tensor(1.0, dtype=torch.float64)

        def forward(self, x1, x2, diag=False, **params):
            res = self.base_kernel.forward(x1, x2, diag=diag, **params)
            if diag and res.dim() == 2:
                res = res.diagonal()
            return torch.as_tensor(self.outputscale, dtype=res.dtype) * res

    class ZeroMean(torch.nn.Module):
        def forward(self, x):
            return torch.zeros(x.shape[:-1], dtype=x.dtype)

        __call__ = forward

    kernels = types.ModuleType("gpytorch.kernels")
    kernels.Kernel, kernels.RBFKernel, kernels.ScaleKernel, kernels.AdditiveKernel = Kernel, RBFKernel, ScaleKernel, AdditiveKernel
    kernels.kernel = Kernel
    means = types.ModuleType("gpytorch.means")
    means.ZeroMean, means.Mean = ZeroMean, torch.nn.Module
    settings = types.ModuleType("gpytorch.settings")
    settings.fast_pred_var = lambda *a, **k: contextlib.nullcontext()
    settings.debug = lambda *a, **k: contextlib.nullcontext()
    models = types.ModuleType("gpytorch.models")
    models.ExactGP = torch.nn.Module
    g.kernels, g.means, g.settings, g.models = kernels, means, settings, models
    for name, mod in (("gpytorch", g), ("gpytorch.kernels", kernels), ("gpytorch.means", means),
                      ("gpytorch.settings", settings), ("gpytorch.models", models)):
        sys.modules[name] = mod
    return g


def _wiener_rbf_kernel(g, WienerKernel, os_w, os_r, ls, dims_rbf):
    kw = WienerKernel(active_dims=[0])
    kr = g.kernels.RBFKernel(ard_num_dims=dims_rbf, active_dims=list(range(1, dims_rbf + 1)))
    k = g.kernels.ScaleKernel(kw) + g.kernels.ScaleKernel(kr)
    k.kernels[0].outputscale = torch.tensor(os_w, dtype=torch.float64)
    k.kernels[1].outputscale = torch.tensor(os_r, dtype=torch.float64)
    k.kernels[1].base_kernel.lengthscale = torch.tensor(ls, dtype=torch.float64).reshape(1, -1)
    return k


def main():
    g = _install_stub()
    sys.path.insert(0, REF)
    from src.gp.recursive_gp import RecursiveGP                       # noqa: E402  (reference code)
    from src.gp.spatiotemporal_gp import ApproxSpatioTemporalGP       # noqa: E402
    from src.gp.wiener_kernel import WienerKernel                     # noqa: E402
    from src.gp.wiener_kernel_temporal import WienerTemporalKernel    # noqa: E402
    torch.set_default_dtype(torch.float64)
    out = {}

    # ---- G1: WienerKernel.forward verbatim
    rng = np.random.default_rng(101)
    t1 = np.sort(rng.uniform(0, 120, (23, 1)), axis=0)
    t2 = rng.uniform(0, 120, (17, 1))
    wk = WienerKernel()
    out["wiener_t1"], out["wiener_t2"] = t1, t2
    out["wiener_cross"] = wk.forward(torch.tensor(t1), torch.tensor(t2)).numpy()
    out["wiener_self"] = wk.forward(torch.tensor(t1), torch.tensor(t1)).numpy()
    out["wiener_diag"] = wk(torch.tensor(t1), torch.tensor(t1), diag=True).numpy()

    # ---- G2: recursive GP == exact GP, isotropic Scale*RBF, duplicates in the training set, full covariance
    #          (tests/gp/test_recursive_gp.py:195-232: s = 3, l = 2, noise 3)
    rng = np.random.default_rng(202)
    x_base = rng.uniform(-5, 5, (20, 3))
    idx = rng.integers(0, 16, 50)
    xt = x_base[idx]
    yt = np.cos(0.5 * np.pi * xt).sum(axis=1)
    xq = rng.uniform(-5, 5, (50, 3))
    k = g.kernels.ScaleKernel(g.kernels.RBFKernel(ard_num_dims=3))
    k.outputscale = torch.tensor(3.0)
    k.base_kernel.lengthscale = torch.tensor(2.0)
    rgp = RecursiveGP(torch.tensor(x_base), Y=None, kernel=k, noise_var=3.0)
    rgp.update(xt, yt)
    yq, cq = rgp.predict(xq, True)
    out.update(rgp_rbf_xt=xt, rgp_rbf_yt=yt, rgp_rbf_xq=xq, rgp_rbf_mean=yq, rgp_rbf_cov=cq)

    # ---- G3: recursive GP with the BattGP kernel structure Scale(Wiener[0]) + Scale(RBF-ARD[1..3]), base == train
    rng = np.random.default_rng(303)
    n = 40
    tt = np.sort(rng.uniform(0.5, 10.0, n))
    ss = rng.uniform(-5, 5, (n, 3))
    xt = np.hstack([tt[:, None], ss])
    yt = np.cos(0.5 * np.pi * ss).sum(axis=1) - np.cos(2 * np.pi * tt / 40)
    xq = np.hstack([rng.uniform(0.5, 10, (30, 1)), rng.uniform(-5, 5, (30, 3))])
    kern = _wiener_rbf_kernel(g, WienerKernel, 10.0, 3.0, [2.0, 3.0, 1.5], 3)
    rgp = RecursiveGP(torch.tensor(xt), Y=None, kernel=kern, noise_var=0.1)
    rgp.update(xt, yt)
    yq, vq = rgp.predict(xq, False)
    out.update(rgp_wr_xt=xt, rgp_wr_yt=yt, rgp_wr_xq=xq, rgp_wr_mean=yq, rgp_wr_var=vq,
               rgp_wr_theta=np.array([10.0, 3.0, 2.0, 3.0, 1.5, 0.1]))

    # ---- G4: Kalman-form spatio-temporal GP == exact Wiener+RBF GP after each of 10 time steps
    #          (tests/gp/test_spatiotemporal_gp.py:218-282: s_w = 10, s_r = 3, l = 2, noise 0.1)
    rng = np.random.default_rng(404)
    tt = np.unique(rng.uniform(0.0, 10.0, 10))
    s_base = rng.uniform(-5, 5, (20, 3))
    st = s_base[rng.choice(20, len(tt))]
    yt = np.cos(0.5 * np.pi * st).sum(axis=1) - np.cos(2 * np.pi * tt / 40)
    sq = rng.uniform(-5, 5, (50, 3))
    ks = g.kernels.ScaleKernel(g.kernels.RBFKernel(ard_num_dims=3))
    ks.outputscale = torch.tensor(3.0)
    ks.base_kernel.lengthscale = torch.tensor(2.0)
    stgp = ApproxSpatioTemporalGP(torch.tensor(s_base), ks, WienerTemporalKernel(outputscale=10.0), 0.1)
    means, vars_ = [], []
    for i in range(len(tt)):
        stgp.time_step(tt[i] - stgp.t)
        stgp.update(st[[i], :], yt[[i]])
        m, v = stgp.predict(sq, full_cov=False)
        means.append(np.asarray(m).reshape(-1))
        vars_.append(np.asarray(v).reshape(-1))
    out.update(stgp_t=tt, stgp_s=st, stgp_y=yt, stgp_sq=sq, stgp_mean=np.array(means), stgp_var=np.array(vars_))

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
