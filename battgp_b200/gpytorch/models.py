"""gpytorch.models.ExactGP backed by the CUDA engine.

Replaces GPyTorch's ExactGP.__call__ + DefaultPredictionStrategy (reached from
/root/reference/src/batt_models/battcellgp_full.py:173 and /root/reference/src/gp/standard_models.py:41):
  eval mode : (cached) fit on the training data -> K_*N alpha for the mean; the variance / covariance solve
              V = K_*N L^-T runs lazily when ``.variance`` / ``._covar`` is read.
  train mode: returns the prior over the training inputs (consumed by ExactMarginalLogLikelihood).
The exact Cholesky path is always taken (DESIGN.md "Semantics vs GPyTorch defaults").
"""
from __future__ import annotations

import math
import warnings
from typing import Optional

import torch

from . import settings
from .. import engine as E
from .distributions import MultivariateNormal
from .kernels import Kernel, LazyKernelMatrix, bind_spec, compute_device, dense_cov, dense_cov_diag, _stage
from .likelihoods import GaussianLikelihood
from .module import Module
from .utils.warnings import GPInputWarning


class GP(Module):
    pass


class PosteriorCovariance:
    """Lazy posterior covariance of the latent f at the query points: k** - V V^T, V = K_*N L^-T (no noise added --
    recursive_gp.py:120 "add no noise_var to be consistent with gpytorch")."""

    def __init__(self, model, state: E.FitState, kernel: Kernel, xq: torch.Tensor, Kq: torch.Tensor, out_dtype, out_device,
                 V: Optional[torch.Tensor] = None):
        self.model, self.state, self.kernel, self.xq, self.Kq = model, state, kernel, xq, Kq
        self.out_dtype, self.out_device = out_dtype, out_device
        self._V: Optional[torch.Tensor] = V      # already solved when the query rows rode through the factorisation

    @property
    def shape(self):
        m = self.xq.shape[0]
        return torch.Size([m, m])

    def _solve(self):
        if self._V is None:
            eng = E.get_engine(self.state.x.device)
            self._V = eng.trsm_rlt(self.state.L, self.state.dinv, self.Kq)      # in place: Kq -> V
            self.Kq = None
        return self._V

    def diagonal(self, *a, **k):
        eng = E.get_engine(self.state.x.device)
        V = self._solve()
        kd = dense_cov_diag(self.kernel, self.xq).detach().to(device=V.device, dtype=torch.float64)
        _, var = eng.predict_tail(V=V, kdiag=kd.contiguous(), min_var=-math.inf)   # clamping happens in .variance
        return var.to(device=self.out_device, dtype=self.out_dtype)

    diag = diagonal

    def to_dense(self):
        eng = E.get_engine(self.state.x.device)
        V = self._solve()
        m = self.xq.shape[0]
        C = E.alloc_matrix(m, m, V.device)
        C.copy_(dense_cov(self.kernel, self.xq, self.xq))
        eng.gemm_nt(V, V, C, alpha=-1.0, beta=1.0)
        return C.to(device=self.out_device, dtype=self.out_dtype)

    evaluate = to_dense

    def detach(self):
        return self.to_dense().detach()

    def cpu(self):
        return self.to_dense().cpu()

    def numpy(self):
        return self.to_dense().cpu().numpy()

    def add_diagonal(self, noise):
        """likelihood(model(x)): predictive covariance of y = f + eps (dense; small M x M)."""
        c = self.to_dense()
        return c + torch.diag_embed(noise.to(c).expand(c.shape[-1]))


class ShardedPosteriorCovariance:
    """Posterior covariance of a GP that was factorised by battgp_b200.sharded over several GPUs: the variance is already
    reduced over the ranks; the dense M x M covariance is not formed in this mode."""

    def __init__(self, var: torch.Tensor):
        self._var = var

    @property
    def shape(self):
        m = self._var.shape[0]
        return torch.Size([m, m])

    def diagonal(self, *a, **k):
        return self._var

    diag = diagonal

    def to_dense(self):
        raise NotImplementedError("the full predictive covariance is not available for a GP sharded over several GPUs "
                                  "(MultiDeviceKernel with n_devices > 1); read .mean / .variance")

    evaluate = to_dense


def _sharded_world(kernel) -> int:
    """cell_gp.py:38-43: BatteryCellGP(n_devices > 1) wraps its kernel in MultiDeviceKernel.  GPyTorch scatters kernel blocks
    with DataParallel inside ONE process; this engine runs one process per GPU, so the knob maps to the block-row-sharded
    factorisation (battgp_b200/sharded.py) when the program was launched with torchrun and the process group spans
    ``n_devices`` ranks.  Returns that world size, or 0 for the single-GPU path."""
    from .kernels import MultiDeviceKernel
    if not isinstance(kernel, MultiDeviceKernel) or len(kernel.device_ids) <= 1:
        return 0
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_world_size() != len(kernel.device_ids):
            raise RuntimeError(f"MultiDeviceKernel(device_ids={kernel.device_ids}) asks for {len(kernel.device_ids)} GPUs but the process "
                               f"group has {dist.get_world_size()} ranks: launch one process per GPU (torchrun --nproc-per-node "
                               f"{len(kernel.device_ids)}), see INTEGRATION.md 'One GP over several GPUs'")
        return dist.get_world_size()
    warnings.warn(f"MultiDeviceKernel over {len(kernel.device_ids)} devices: battgp_b200 shards ONE GP across GPUs with one process "
                  "per GPU (torchrun + battgp_b200.sharded, INTEGRATION.md); in this single-process run the factorisation stays on "
                  "one GPU (same results, single-GPU memory limit)", RuntimeWarning)
    return 0


class ExactGP(GP):
    def __init__(self, train_inputs, train_targets, likelihood):
        if train_inputs is not None and torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        if train_inputs is not None and not all(torch.is_tensor(t) for t in train_inputs):
            raise RuntimeError("Train inputs must be a tensor, or a list/tuple of tensors")
        if not isinstance(likelihood, GaussianLikelihood):
            raise RuntimeError("ExactGP can only handle Gaussian likelihoods")
        super().__init__()
        self.train_inputs = None if train_inputs is None else tuple(t.unsqueeze(-1) if t.dim() == 1 else t for t in train_inputs)
        self.train_targets = train_targets
        self.likelihood = likelihood
        self.prediction_strategy = None       # (signature, FitState, kernel)

    # nn.Module.to()/cuda()/double() must also move the training data (GPyTorch's ExactGP._apply)
    def _apply(self, fn, *a, **k):
        if self.train_inputs is not None:
            self.train_inputs = tuple(fn(t) for t in self.train_inputs)
            self.train_targets = fn(self.train_targets)
        self.prediction_strategy = None
        return super()._apply(fn, *a, **k)

    def train(self, mode=True):
        if mode:
            self.prediction_strategy = None
        return super().train(mode)

    def set_train_data(self, inputs=None, targets=None, strict=True):
        if inputs is not None:
            if torch.is_tensor(inputs):
                inputs = (inputs,)
            self.train_inputs = tuple(t.unsqueeze(-1) if t.dim() == 1 else t for t in inputs)
        if targets is not None:
            self.train_targets = targets
        self.prediction_strategy = None

    def forward(self, *x):
        raise NotImplementedError

    # ------------------------------------------------------------------------------------------------
    def _signature(self):
        ps = [p.detach().reshape(-1).to(torch.float64) for p in self.parameters()]
        return tuple(torch.cat(ps).tolist()) if ps else ()

    def _fit_state(self, xq=None):
        """(Re)factorise when no valid cache exists.  ``xq``: query points of the call that triggers the fit -- their
        cross-covariance rows are appended to K and solved inside the factorisation (E.fit(..., xq=...))."""
        sig = self._signature()
        if self.prediction_strategy is not None and self.prediction_strategy[0] == sig:
            return self.prediction_strategy[1], self.prediction_strategy[2]
        train_x = self.train_inputs[0]
        prior = self.forward(*self.train_inputs)
        cov = prior._covar
        if not isinstance(cov, LazyKernelMatrix):
            raise RuntimeError("ExactGP.forward must return MultivariateNormal(mean, self.covar_module(x))")
        kernel = cov.kernel
        dev = compute_device(train_x)
        x64 = _stage(train_x, dev)
        resid = _stage(self.train_targets - prior.mean, dev)
        noise = float(self.likelihood.noise.detach().reshape(-1)[0])
        binding = bind_spec(kernel, train_x.shape[-1])
        if binding is not None:
            st = E.fit(binding.to_spec(), x64, resid, noise, xq=None if xq is None else _stage(xq, dev))
        else:
            def kbuilder(out, nz):
                k = dense_cov(kernel, train_x, train_x).to(device=dev, dtype=torch.float64)
                out.copy_(k)
                out.diagonal().add_(nz)
            st = E.fit(E.KernelSpec([]), x64, resid, noise, kbuilder=kbuilder)
        self.prediction_strategy = (sig, st, kernel)
        return st, kernel

    def _sharded_call(self, kernel, xq, inputs, kwargs):
        """Eval-mode call of a model whose kernel is MultiDeviceKernel(n_devices = world size): every rank runs this same call
        (SPMD) with the same data; the N x N factorisation is block-row-sharded over the ranks (battgp_b200/sharded.py)."""
        from ..sharded import ShardedGP
        train_x = self.train_inputs[0]
        binding = bind_spec(kernel, train_x.shape[-1])
        if binding is None:
            raise RuntimeError("MultiDeviceKernel over several GPUs needs a kernel the engine builds natively "
                               "(Wiener / RBF / Matern-5/2 / Periodic under ScaleKernel, summed)")
        dev = compute_device(train_x)
        sig = self._signature()
        if self.prediction_strategy is None or self.prediction_strategy[0] != ("sharded", sig):
            prior = self.forward(*self.train_inputs)
            resid = _stage(self.train_targets - prior.mean, dev)
            noise = float(self.likelihood.noise.detach().reshape(-1)[0])
            n = train_x.shape[0]
            gp = ShardedGP(binding.to_spec(), _stage(train_x, dev), resid, noise, nb=2048 if n >= 100000 else 1024)
            gp.fit(_stage(xq, dev))          # the first query grid rides through the factorisation (fit once, predict once: battgp_full.py:98-120)
            self.prediction_strategy = (("sharded", sig), gp, kernel)
        gp = self.prediction_strategy[1]
        test_prior = self.forward(*inputs, **kwargs)
        mean, var = gp.predict(_stage(xq, dev), clamp=False)
        mean = mean.to(device=xq.device, dtype=xq.dtype) + test_prior.mean
        return MultivariateNormal(mean, ShardedPosteriorCovariance(var.to(device=xq.device, dtype=xq.dtype)))

    def __call__(self, *args, **kwargs):
        inputs = [a.unsqueeze(-1) if a.dim() == 1 else a for a in args]
        if self.training:
            if self.train_inputs is None:
                raise RuntimeError("train_inputs, train_targets cannot be None in training mode. Call .eval() for prior "
                                   "predictions, or call .set_train_data() to add training data.")
            if settings.debug.on():
                if not all(torch.equal(a, b) for a, b in zip(self.train_inputs, inputs)):
                    raise RuntimeError("You must train on the training inputs!")
            return self.forward(*inputs, **kwargs)
        if self.train_inputs is None or self.train_targets is None:
            return self.forward(*inputs, **kwargs)       # prior mode
        if settings.debug.on() and all(a.shape == b.shape and torch.equal(a, b) for a, b in zip(self.train_inputs, inputs)):
            warnings.warn("The input matches the stored training data. Did you forget to call model.train()?", GPInputWarning)
        xq = inputs[0]
        prior_cov = self.forward(*self.train_inputs)._covar
        if isinstance(prior_cov, LazyKernelMatrix) and _sharded_world(prior_cov.kernel):
            return self._sharded_call(prior_cov.kernel, xq, inputs, kwargs)
        st, kernel = self._fit_state(xq)
        test_prior = self.forward(*inputs, **kwargs)
        dev = st.x.device
        eng = E.get_engine(dev)
        binding = bind_spec(kernel, xq.shape[-1])
        xq64 = _stage(xq, dev)
        if binding is not None:
            Kq = eng.cov_build(binding.to_spec(), xq64, st.x)
        else:
            Kq = E.alloc_matrix(xq.shape[0], st.x.shape[0], dev)
            Kq.copy_(dense_cov(kernel, xq, self.train_inputs[0]).to(device=dev, dtype=torch.float64))
        mean, _ = eng.predict_tail(Kq=Kq, alpha=st.alpha)
        mean = mean.to(device=xq.device, dtype=xq.dtype) + test_prior.mean
        V = None
        if st.V is not None and st.xq is not None and st.xq.shape == xq64.shape and torch.equal(st.xq, xq64):
            V = st.V                                 # this call triggered the fit: the solve is already done
        return MultivariateNormal(mean, PosteriorCovariance(self, st, kernel, xq, Kq, xq.dtype, xq.device, V=V))
