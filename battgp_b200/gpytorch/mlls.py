"""gpytorch.mlls.ExactMarginalLogLikelihood (training.py:27,40,79,126): mll = LML / N with an analytic backward.

forward : fused covariance build + Cholesky + alpha + logdet on the GPU (one host sync to read the hyper-parameters)
backward: K^-1 by POTRI on the DMMA pipe, then ONE fused pass that recomputes dK/dtheta tile-wise and reduces it against
          (alpha alpha^T - K^-1)  (bgp_lml_grad) -- replaces autograd through torch.linalg.cholesky (training.py:41,140).
Gradients reach the raw parameters through the constraint transforms by ordinary autograd.
"""
from __future__ import annotations

import torch

from .. import engine as E
from .distributions import MultivariateNormal
from .kernels import LazyKernelMatrix, bind_spec, compute_device, dense_cov, torch_cov, _stage
from .module import Module


class _NativeLML(torch.autograd.Function):
    @staticmethod
    def forward(ctx, binding, x64, resid, noise, *params):
        n = x64.shape[0]
        flat = torch.cat([noise.detach().reshape(-1)[:1].to(torch.float64)] +
                         [p.detach().reshape(-1).to(torch.float64) for p in params]).tolist()
        values, o = [], 1
        for p in params:
            values.append(flat[o:o + p.numel()])
            o += p.numel()
        spec = binding.to_spec(values)
        st = E.fit(spec, x64, _stage(resid, x64.device), flat[0])
        ctx.st, ctx.binding, ctx.n = st, binding, n
        ctx.meta = (resid.dtype, resid.device, noise.dtype, noise.device, noise.shape, [(p.dtype, p.device) for p in params])
        ctx.inverted = False
        # GPyTorch promotes: BatteryCellGP keeps fp32 hyper-parameters next to fp64 data and its loss is fp64
        return torch.tensor(st.lml / n, dtype=torch.promote_types(noise.dtype, resid.dtype), device=noise.device)

    @staticmethod
    def backward(ctx, gout):
        st, n = ctx.st, ctx.n
        eng = E.get_engine(st.x.device)
        if not ctx.inverted:
            eng.potri(st.L, st.dinv)          # L -> K^-1 in place (the factor is not needed again in training mode)
            ctx.inverted = True
        g = eng.lml_grad(st.spec, st.noise + st.jitter, st.x, st.L, st.alpha) / n
        rd, rdev, nd, ndev, nshape, pmeta = ctx.meta
        go = gout.to(torch.float64)
        g_noise = (g[0] * go.to(g.device)).to(device=ndev, dtype=nd).reshape(1).expand(nshape).clone()
        g_resid = (-st.alpha / n * go.to(st.alpha.device)).to(device=rdev, dtype=rd) if ctx.needs_input_grad[2] else None
        routed = ctx.binding.route_grads(g[1:])
        g_params = tuple((r * go.to(r.device)).to(device=dev, dtype=dt) for r, (dt, dev) in zip(routed, pmeta))
        return (None, None, g_resid, g_noise) + g_params


class _DenseLML(torch.autograd.Function):
    """Generic kernels: K comes from the user's own torch forward; dLML/dK = 0.5 (alpha alpha^T - K^-1) goes back into it."""

    @staticmethod
    def forward(ctx, K, resid):
        dev = compute_device(K)
        n = K.shape[0]
        Kd = K.detach().to(device=dev, dtype=torch.float64)

        def kbuilder(out, jitter):
            out.copy_(Kd)
            if jitter:
                out.diagonal().add_(jitter)
        st = E.fit(E.KernelSpec([]), torch.empty(n, 1, dtype=torch.float64, device=dev), _stage(resid, dev), 0.0, kbuilder=kbuilder)
        ctx.st, ctx.n = st, n
        ctx.meta = (K.dtype, K.device, resid.dtype, resid.device)
        return torch.tensor(st.lml / n, dtype=torch.promote_types(K.dtype, resid.dtype), device=K.device)

    @staticmethod
    def backward(ctx, gout):
        st, n = ctx.st, ctx.n
        eng = E.get_engine(st.L.device)
        eng.potri(st.L, st.dinv)
        kinv = torch.tril(st.L)
        kinv = kinv + torch.tril(kinv, -1).T
        kd, kdev, rd, rdev = ctx.meta
        go = float(gout)
        gK = (0.5 * go / n) * (torch.outer(st.alpha, st.alpha) - kinv)
        g_resid = (-go / n) * st.alpha
        return gK.to(device=kdev, dtype=kd), g_resid.to(device=rdev, dtype=rd)


class MarginalLogLikelihood(Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model


class ExactMarginalLogLikelihood(MarginalLogLikelihood):
    def forward(self, function_dist, target, *params, **kwargs):
        if not isinstance(function_dist, MultivariateNormal):
            raise RuntimeError("ExactMarginalLogLikelihood can only operate on Gaussian random variables")
        cov = function_dist._covar
        resid = target - function_dist.mean
        noise = self.likelihood.noise
        if isinstance(cov, LazyKernelMatrix):
            x = cov.x1
            binding = bind_spec(cov.kernel, x.shape[-1])
            if binding is not None:
                dev = compute_device(x)
                return _NativeLML.apply(binding, _stage(x, dev), resid, noise, *binding.param_tensors())
            K = torch_cov(cov.kernel, x, x) if torch.is_grad_enabled() else dense_cov(cov.kernel, x, x)
        else:
            # e.g. the eval-mode posterior handed back to the mll (training.py:100-103 does exactly that after fitting):
            # log N(y; posterior mean, posterior covariance + noise), like GPyTorch's likelihood(function_dist)
            K = cov if torch.is_tensor(cov) else cov.to_dense()
        K = K + torch.diag_embed(noise.to(K).expand(K.shape[-1]))
        return _DenseLML.apply(K, resid)

    __call__ = torch.nn.Module.__call__
