"""gpytorch.kernels surface used by BattGP backed by the fused CUDA covariance build.

``Kernel.__call__`` returns a lazy ``LazyKernelMatrix`` (nothing is computed until a consumer asks), ``Kernel.forward``
returns a dense tensor like GPyTorch does -- callers such as /root/reference/src/gp/recursive_gp.py:52-57 and
spatiotemporal_gp.py:157-162 use it directly.  For kernel trees made of the native terms (RBF, Matern-5/2, Periodic and
any registered class such as the reference's ``WienerKernel``; optionally wrapped in ScaleKernel / summed) the whole
tree is evaluated by ONE launch of ``bgp_cov_build``.  Unknown user kernels run their own torch ``forward`` and only the
factorisation / solves go through the engine.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch

from .. import _lib
from .. import engine as E
from .constraints import Positive
from .module import Module

# class name -> native term type; the reference's src/gp/wiener_kernel.py:WienerKernel is picked up by name
NATIVE_KERNELS = {"WienerKernel": _lib.WIENER}


def register_native_kernel(cls_or_name, term_type: int) -> None:
    NATIVE_KERNELS[cls_or_name if isinstance(cls_or_name, str) else cls_or_name.__name__] = term_type


def compute_device(t: torch.Tensor) -> torch.device:
    """Where the arithmetic happens: the tensor's own GPU, else the current GPU (CPU tensors are staged through it).
    There is no CPU implementation of the path -- fail loudly when no GPU is present."""
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise _lib.BattGPLibraryError("battgp_b200: no CUDA device available and there is no CPU fallback for the GP path")
    return torch.device("cuda", torch.cuda.current_device())


def _stage(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float64).contiguous()


class Kernel(Module):
    has_lengthscale = False
    is_stationary = True

    def __init__(self, ard_num_dims: Optional[int] = None, batch_shape=torch.Size(), active_dims=None,
                 lengthscale_prior=None, lengthscale_constraint=None, eps: float = 1e-6, **kwargs):
        super().__init__()
        if active_dims is not None and not torch.is_tensor(active_dims):
            active_dims = torch.tensor(list(active_dims), dtype=torch.long)
        self.register_buffer("active_dims", active_dims)
        self.ard_num_dims = ard_num_dims
        self.batch_shape = batch_shape
        self.eps = eps
        if self.has_lengthscale:
            n = 1 if ard_num_dims is None else ard_num_dims
            self.register_parameter("raw_lengthscale", torch.nn.Parameter(torch.zeros(*batch_shape, 1, n)))
            self.register_constraint("raw_lengthscale", lengthscale_constraint if lengthscale_constraint is not None else Positive())

    # -- hyper-parameters
    @property
    def lengthscale(self):
        return self._get_constrained("raw_lengthscale") if self.has_lengthscale else None

    @lengthscale.setter
    def lengthscale(self, value):
        if not self.has_lengthscale:
            raise RuntimeError("Kernel has no lengthscale.")
        self._set_constrained("raw_lengthscale", value)

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        for b in self.buffers():
            if b is not None:
                return b.device
        return torch.device("cpu")

    @property
    def dtype(self):
        for p in self.parameters():
            return p.dtype
        return torch.get_default_dtype()

    # -- evaluation
    def covar_dist(self, x1, x2, diag=False, last_dim_is_batch=False, square_dist=False, **params):
        """Euclidean (squared) distance; kept for user kernels such as wiener_kernel.py:11."""
        if diag:
            d2 = (x1 - x2).pow(2).sum(-1)
            return d2 if square_dist else d2.clamp_min(1e-30).sqrt()
        d = torch.cdist(x1, x2)
        return d.pow(2) if square_dist else d

    def forward(self, x1, x2, diag=False, **params):
        raise NotImplementedError

    def _select(self, x):
        if x.dim() == 1:
            x = x.unsqueeze(-1)
        if self.active_dims is not None:
            x = x.index_select(-1, self.active_dims.to(x.device))
        return x

    def __call__(self, x1, x2=None, diag=False, last_dim_is_batch=False, **params):
        x1_ = x1.unsqueeze(-1) if x1.dim() == 1 else x1
        x2_ = x1_ if x2 is None else (x2.unsqueeze(-1) if x2.dim() == 1 else x2)
        if diag:
            res = self.forward(self._select(x1_), self._select(x2_), diag=True, **params)
            if res.dim() == 2 and res.shape[0] == res.shape[1]:      # GPyTorch's fallback
                res = res.diagonal()
            return res
        return LazyKernelMatrix(self, x1_, x2_, params)

    def __add__(self, other):
        ks = (list(self.kernels) if isinstance(self, AdditiveKernel) else [self]) + \
             (list(other.kernels) if isinstance(other, AdditiveKernel) else [other])
        return AdditiveKernel(*ks)

    def __getitem__(self, index):
        return self

    def num_outputs_per_input(self, x1, x2):
        return 1


# ------------------------------------------------------------------------------------------ native spec extraction
@dataclass
class _BoundTerm:
    type: int
    dims: List[int]
    outputscale: Optional[torch.Tensor]     # None -> 1.0 (bare kernel without ScaleKernel)
    lengthscale: Optional[torch.Tensor]     # [.., 1, D] or [.., 1, 1]
    period: Optional[torch.Tensor]


class SpecBinding:
    """A kernel tree flattened to native terms, holding references to the (constrained) hyper-parameter tensors so
    that values can be read for the CUDA build and gradients routed back to them."""

    def __init__(self, terms: List[_BoundTerm]):
        self.terms = terms

    def param_tensors(self) -> List[torch.Tensor]:
        out = []
        for t in self.terms:
            for p in (t.outputscale, t.lengthscale, t.period):
                if p is not None:
                    out.append(p)
        return out

    def to_spec(self, values: Optional[List[List[float]]] = None) -> E.KernelSpec:
        """values: per param_tensors() entry, flattened floats (one host sync for all of them)."""
        if values is None:
            ps = self.param_tensors()
            if ps:
                flat = torch.cat([p.detach().reshape(-1).to(torch.float64) for p in ps]).tolist()
            else:
                flat = []
            values, o = [], 0
            for p in ps:
                values.append(flat[o:o + p.numel()])
                o += p.numel()
        it = iter(values)
        terms = []
        for t in self.terms:
            os_ = next(it)[0] if t.outputscale is not None else 1.0
            ls = list(next(it)) if t.lengthscale is not None else []
            per = list(next(it)) if t.period is not None else []
            nd = len(t.dims)
            if len(ls) == 1 and nd > 1:
                ls = ls * nd
            if len(per) == 1 and nd > 1:
                per = per * nd
            terms.append(E.Term(t.type, t.dims, os_, tuple(ls), tuple(per)))
        return E.KernelSpec(terms)

    def route_grads(self, g: torch.Tensor) -> List[torch.Tensor]:
        """g: slots from bgp_lml_grad WITHOUT the leading noise slot.  Returns one gradient per param_tensors() entry."""
        out, o = [], 0
        for t in self.terms:
            nd = len(t.dims)
            g_os = g[o]; o += 1
            if t.outputscale is not None:
                out.append(g_os.reshape(t.outputscale.shape))
            if t.type != _lib.WIENER:
                g_ls = g[o:o + nd]; o += nd
                if t.lengthscale is not None:
                    out.append((g_ls.sum() if t.lengthscale.numel() == 1 else g_ls).reshape(t.lengthscale.shape))
                if t.type == _lib.PERIODIC:
                    g_p = g[o:o + nd]; o += nd
                    if t.period is not None:
                        out.append((g_p.sum() if t.period.numel() == 1 else g_p).reshape(t.period.shape))
        return out


def _native_type(k: "Kernel") -> Optional[int]:
    if isinstance(k, RBFKernel):
        return _lib.RBF
    if isinstance(k, MaternKernel):
        return _lib.MATERN52 if abs(k.nu - 2.5) < 1e-12 else None
    if isinstance(k, PeriodicKernel):
        return _lib.PERIODIC
    for cls in type(k).__mro__:
        if cls.__name__ in NATIVE_KERNELS:
            return NATIVE_KERNELS[cls.__name__]
    return None


def bind_spec(kernel: "Kernel", d_total: int, outer_dims: Optional[List[int]] = None) -> Optional[SpecBinding]:
    """Flatten ``kernel`` into native terms over the columns of a ``d_total``-column input, or None if some node is
    not native (the caller then falls back to the kernel's own torch forward for the *build* only)."""
    if isinstance(kernel, MultiDeviceKernel):
        return bind_spec(kernel.base_kernel, d_total, outer_dims)
    cols = list(range(d_total)) if outer_dims is None else outer_dims
    if kernel.active_dims is not None:
        cols = [cols[int(i)] for i in kernel.active_dims.tolist()]
    if isinstance(kernel, AdditiveKernel):
        terms = []
        for k in kernel.kernels:
            b = bind_spec(k, d_total, cols)
            if b is None:
                return None
            terms.extend(b.terms)
        return SpecBinding(terms) if len(terms) <= _lib.MAX_TERMS else None
    os_ = None
    base = kernel
    if isinstance(kernel, ScaleKernel):
        os_ = kernel.outputscale
        base = kernel.base_kernel
        if isinstance(base, (ScaleKernel, AdditiveKernel, MultiDeviceKernel)):
            return None
        # ScaleKernel copies active_dims from its base kernel: columns were already selected above
    ty = _native_type(base)
    if ty is None or len(cols) > _lib.MAX_DIMS or len(cols) < 1:
        return None
    if ty == _lib.WIENER and len(cols) != 1:
        return None
    ls = base.lengthscale if base.has_lengthscale else None
    if ty != _lib.WIENER and ls is None:
        return None
    if ls is not None and ls.numel() not in (1, len(cols)):
        return None
    per = base.period_length if ty == _lib.PERIODIC else None
    if per is not None and per.numel() not in (1, len(cols)):
        return None
    return SpecBinding([_BoundTerm(ty, cols, os_, ls, per)])


def dense_cov(kernel: "Kernel", x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """k(x1, x2) as a dense tensor with x1's dtype/device; arithmetic in fp64 on the GPU."""
    b = bind_spec(kernel, x1.shape[-1])
    if b is None:
        xs1, xs2 = kernel._select(x1), kernel._select(x2)
        if isinstance(kernel, InducingPointKernel):     # fp64 all the way into the exact path (callers widen anyway)
            return kernel._forward64(xs1, xs2)
        return kernel.forward(xs1, xs2)
    dev = compute_device(x1)
    out = E.get_engine(dev).cov_build(b.to_spec(), _stage(x1, dev), _stage(x2, dev))
    return out.to(device=x1.device, dtype=x1.dtype)


def torch_cov(kernel: "Kernel", x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """Differentiable dense k(x1, x2) in plain torch ops.  Only used for the LML of kernel trees that contain a
    NON-native user kernel (autograd has to flow through the user's forward, so the whole tree is evaluated in torch on
    the inputs' device); BattGP's own kernels never take this route."""
    if isinstance(kernel, MultiDeviceKernel):
        return torch_cov(kernel.base_kernel, x1, x2)
    if isinstance(kernel, AdditiveKernel):
        a, b = kernel._select(x1), kernel._select(x2)
        out = 0
        for k in kernel.kernels:
            out = out + torch_cov(k, a, b)
        return out
    a, b = kernel._select(x1), kernel._select(x2)
    if isinstance(kernel, ScaleKernel):
        base = kernel.base_kernel
        return kernel.outputscale.to(a) * _torch_base(base, a, b)
    return _torch_base(kernel, a, b)


def _torch_base(k: "Kernel", a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    ty = _native_type(k)
    if ty == _lib.RBF or ty == _lib.MATERN52:
        ls = k.lengthscale.to(a).reshape(1, -1)
        d2 = ((a / ls).unsqueeze(-2) - (b / ls).unsqueeze(-3)).pow(2).sum(-1)
        if ty == _lib.RBF:
            return torch.exp(-0.5 * d2)
        r = d2.clamp_min(1e-30).sqrt()
        return (1.0 + math.sqrt(5.0) * r + (5.0 / 3.0) * r * r) * torch.exp(-math.sqrt(5.0) * r)
    if ty == _lib.PERIODIC:
        ls = k.lengthscale.to(a).reshape(1, -1)
        p = k.period_length.to(a).reshape(1, -1)
        diff = (a.unsqueeze(-2) - b.unsqueeze(-3)) * (math.pi / p)
        return torch.exp(-2.0 * (torch.sin(diff).pow(2) / ls).sum(-1))
    return k.forward(a, b)


def dense_cov_diag(kernel: "Kernel", x: torch.Tensor) -> torch.Tensor:
    b = bind_spec(kernel, x.shape[-1])
    if b is None:
        if isinstance(kernel, InducingPointKernel):
            return kernel._forward64(kernel._select(x), kernel._select(x), diag=True)
        return kernel(x, x, diag=True)
    dev = compute_device(x)
    return E.get_engine(dev).cov_diag(b.to_spec(), _stage(x, dev)).to(device=x.device, dtype=x.dtype)


class LazyKernelMatrix:
    """What ``kernel(x1, x2)`` returns: the kernel and its inputs; evaluated on demand (GPyTorch's
    LazyEvaluatedKernelTensor).  ``add_noise`` is set by GaussianLikelihood.__call__."""

    def __init__(self, kernel, x1, x2, params=None, noise: Optional[torch.Tensor] = None):
        self.kernel, self.x1, self.x2, self.params, self.noise = kernel, x1, x2, params or {}, noise

    @property
    def shape(self):
        return torch.Size([self.x1.shape[-2], self.x2.shape[-2]])

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def dtype(self):
        return self.x1.dtype

    @property
    def device(self):
        return self.x1.device

    def to_dense(self) -> torch.Tensor:
        k = dense_cov(self.kernel, self.x1, self.x2)
        if self.noise is not None:
            k = k + torch.diag_embed(self.noise.to(k).expand(k.shape[-1]))
        return k

    evaluate = to_dense

    def evaluate_kernel(self):
        return self

    def diagonal(self, *a, **k):
        d = dense_cov_diag(self.kernel, self.x1) if self.x1 is self.x2 or torch.equal(self.x1, self.x2) else self.to_dense().diagonal()
        return d + self.noise.to(d) if self.noise is not None else d

    diag = diagonal

    def add_diagonal(self, noise):
        return LazyKernelMatrix(self.kernel, self.x1, self.x2, self.params, noise if self.noise is None else self.noise + noise)

    def detach(self):
        return self.to_dense().detach()

    def cpu(self):
        return self.to_dense().cpu()

    def numpy(self):
        return self.to_dense().detach().cpu().numpy()


# ------------------------------------------------------------------------------------------ concrete kernels
class RBFKernel(Kernel):
    """exp(-0.5 * sum_d ((a_d - b_d)/l_d)^2)   (cell_gp.py:33, standard_models.py:24)"""
    has_lengthscale = True

    def forward(self, x1, x2, diag=False, **params):
        if diag:
            ls = self.lengthscale.to(x1)
            return torch.exp(-0.5 * ((x1 - x2) / ls.reshape(1, -1)).pow(2).sum(-1))
        return _native_forward(self, _lib.RBF, x1, x2)


class MaternKernel(Kernel):
    has_lengthscale = True

    def __init__(self, nu: float = 2.5, **kwargs):
        if nu not in (0.5, 1.5, 2.5):
            raise RuntimeError("nu expected to be 0.5, 1.5, or 2.5")
        super().__init__(**kwargs)
        self.nu = nu

    def forward(self, x1, x2, diag=False, **params):
        if self.nu != 2.5:
            raise NotImplementedError("battgp_b200 implements Matern nu=2.5 (BASELINE.json config 3) natively only")
        if diag:
            return torch.ones(x1.shape[:-1], dtype=x1.dtype, device=x1.device)
        return _native_forward(self, _lib.MATERN52, x1, x2)


class PeriodicKernel(Kernel):
    """exp(-2 sum_d sin^2(pi (a_d - b_d)/p_d) / l_d)  -- GPyTorch's parameterisation (divides by l, not l^2)."""
    has_lengthscale = True

    def __init__(self, period_length_prior=None, period_length_constraint=None, **kwargs):
        super().__init__(**kwargs)
        n = 1 if self.ard_num_dims is None else self.ard_num_dims
        self.register_parameter("raw_period_length", torch.nn.Parameter(torch.zeros(*self.batch_shape, 1, n)))
        self.register_constraint("raw_period_length", period_length_constraint if period_length_constraint is not None else Positive())

    @property
    def period_length(self):
        return self._get_constrained("raw_period_length")

    @period_length.setter
    def period_length(self, value):
        self._set_constrained("raw_period_length", value)

    def forward(self, x1, x2, diag=False, **params):
        if diag:
            return torch.ones(x1.shape[:-1], dtype=x1.dtype, device=x1.device)
        return _native_forward(self, _lib.PERIODIC, x1, x2)


def _native_forward(k: Kernel, ty: int, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """Dense k(x1, x2) for a single native base kernel on already-selected columns."""
    d = x1.shape[-1]
    ls = k.lengthscale
    per = k.period_length if ty == _lib.PERIODIC else None
    b = SpecBinding([_BoundTerm(ty, list(range(d)), None, ls, per)])
    dev = compute_device(x1)
    out = E.get_engine(dev).cov_build(b.to_spec(), _stage(x1, dev), _stage(x2, dev))
    return out.to(device=x1.device, dtype=x1.dtype)


class ScaleKernel(Kernel):
    """s * k_base  (cell_gp.py:34-36); copies ``active_dims`` from the base kernel like GPyTorch does."""

    def __init__(self, base_kernel: Kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        kwargs.pop("device", None)       # tests/gp/test_spatiotemporal_gp.py passes device=...; GPyTorch ignores it too
        if base_kernel.active_dims is not None:
            kwargs["active_dims"] = base_kernel.active_dims
        super().__init__(**kwargs)
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", torch.nn.Parameter(torch.zeros(self.batch_shape)))
        self.register_constraint("raw_outputscale", outputscale_constraint if outputscale_constraint is not None else Positive())

    @property
    def outputscale(self):
        return self._get_constrained("raw_outputscale")

    @outputscale.setter
    def outputscale(self, value):
        self._set_constrained("raw_outputscale", value)

    @property
    def is_stationary(self):
        return self.base_kernel.is_stationary

    def forward(self, x1, x2, diag=False, **params):
        ty = _native_type(self.base_kernel)
        if ty is not None and not diag and (ty == _lib.WIENER or self.base_kernel.has_lengthscale):
            d = x1.shape[-1]
            base = self.base_kernel
            b = SpecBinding([_BoundTerm(ty, list(range(d)), self.outputscale, base.lengthscale if base.has_lengthscale else None,
                                        base.period_length if ty == _lib.PERIODIC else None)])
            dev = compute_device(x1)
            out = E.get_engine(dev).cov_build(b.to_spec(), _stage(x1, dev), _stage(x2, dev))
            return out.to(device=x1.device, dtype=x1.dtype)
        res = self.base_kernel.forward(x1, x2, diag=diag, **params)
        if diag and res.dim() == 2 and res.shape[0] == res.shape[1]:
            res = res.diagonal()
        return res * self.outputscale.to(res)


class AdditiveKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = torch.nn.ModuleList(kernels)

    @property
    def is_stationary(self):
        return all(k.is_stationary for k in self.kernels)

    def forward(self, x1, x2, diag=False, **params):
        if not diag:
            b = bind_spec(self, x1.shape[-1])
            if b is not None:
                dev = compute_device(x1)
                out = E.get_engine(dev).cov_build(b.to_spec(), _stage(x1, dev), _stage(x2, dev))
                return out.to(device=x1.device, dtype=x1.dtype)
        out = 0
        for k in self.kernels:
            r = k(x1, x2, diag=diag, **params)
            out = out + (r if torch.is_tensor(r) else r.to_dense())
        return out


class MultiDeviceKernel(Kernel):
    """cell_gp.py:38-43.  GPyTorch scatters kernel blocks over devices with DataParallel inside one process; here ONE GP's
    factorisation is block-row-sharded over one process per GPU (battgp_b200/sharded.py): an ExactGP whose kernel is a
    MultiDeviceKernel with more than one device id takes that path in eval mode when the program runs under torchrun with a
    process group of that many ranks (gpytorch/models.py _sharded_world), and warns + stays on one GPU otherwise.  The
    covariance itself is ``base_kernel``'s."""

    def __init__(self, base_kernel, device_ids, output_device=None, **kwargs):
        super().__init__()
        self.base_kernel = base_kernel
        self.device_ids = list(device_ids)
        self.output_device = output_device

    def forward(self, x1, x2, diag=False, **params):
        return self.base_kernel.forward(self.base_kernel._select(x1), self.base_kernel._select(x2), diag=diag, **params)


class InducingPointKernel(Kernel):
    """standard_models.py:83-87 (SparseScaledRBFModel, SGPR): the subset-of-regressors covariance through m inducing points u,
        k_Q(x1, x2) = K_x1,u K_uu^-1 K_u,x2,
    which is what GPyTorch's InducingPointKernel returns in eval mode (a low-rank root); in training mode the diagonal of the
    train block is corrected to the exact prior variance (FITC-style ``LowRankRootAddedDiagLinearOperator``).  Here the factor
    A = K_xu L_uu^-T is formed with the engine (bgp_potrf, bgp_trsm_rlt, bgp_gemm_nt) and the N x N matrix Q = A A^T is handed to
    the exact path as a dense covariance: the posterior mean / variance equal GPyTorch's SGPR prediction, computed with the
    O(N^3) exact factorisation instead of the O(N m^2) Woodbury form (SGPR is an approximate method outside the hot path,
    SURVEY.md section 2 row 4 -- supported for completeness of the reference surface, not optimised).  The variational trace
    term GPyTorch adds to the MLL for learning the inducing points is not implemented (the reference never trains this
    model: its ``optimize`` calls a function that does not exist)."""

    def __init__(self, base_kernel, inducing_points, likelihood, active_dims=None, **kwargs):
        super().__init__(active_dims=active_dims)
        self.base_kernel = base_kernel
        self.likelihood = likelihood
        if inducing_points.dim() == 1:
            inducing_points = inducing_points.unsqueeze(-1)
        self.register_parameter("inducing_points", torch.nn.Parameter(inducing_points))

    def _root(self, x: torch.Tensor, u: torch.Tensor, Luu: torch.Tensor, dinv: torch.Tensor, dev) -> torch.Tensor:
        """A = K_xu L_uu^-T  [n, m], fp64 on the compute device."""
        eng = E.get_engine(dev)
        Kxu = E.alloc_matrix(x.shape[0], u.shape[0], dev)
        Kxu.copy_(dense_cov(self.base_kernel, x, u).to(device=dev, dtype=torch.float64))
        return eng.trsm_rlt(Luu, dinv, Kxu)

    def _forward64(self, x1, x2, diag=False):
        """k_Q(x1, x2) in fp64 on the compute device.  fp32 inputs (standard_models.py:66-67) are widened FIRST, so neither K_xu
        nor Q is ever rounded to fp32 on its way into the exact path (cond(Q + sigma^2 I) ~ 1e4 would turn that rounding into
        1e-3 relative errors of the posterior mean)."""
        dev = compute_device(x1)
        eng = E.get_engine(dev)
        same = x1 is x2 or (x1.shape == x2.shape and torch.equal(x1, x2))
        x1 = x1.detach().to(torch.float64)
        x2 = x1 if same else x2.detach().to(torch.float64)
        u = self.inducing_points.detach().to(device=x1.device, dtype=torch.float64)
        m = u.shape[0]
        Kuu = E.alloc_matrix(m, m, dev)
        Kuu.copy_(dense_cov(self.base_kernel, u, u).to(device=dev, dtype=torch.float64))
        jitter = 0.0
        for attempt in range(len(E.JITTERS_F64) + 1):          # psd_safe_cholesky of K_uu
            L = Kuu.clone()
            if jitter:
                L.diagonal().add_(jitter)
            info, _, dinv = eng.potrf(L)
            if info == 0:
                break
            if attempt == len(E.JITTERS_F64):
                raise E.NotPSDError("inducing-point covariance K_uu is not positive definite")
            jitter = E.JITTERS_F64[attempt]
        A1 = self._root(x1, u, L, dinv, dev)
        A2 = A1 if same else self._root(x2, u, L, dinv, dev)
        if diag:
            out = torch.empty(x1.shape[0], dtype=torch.float64, device=dev)
            if same:
                eng.rowsumsq(A1, out, False)
            else:
                out = (A1 * A2).sum(-1)
            if self.training and same:
                out = torch.maximum(out, dense_cov_diag(self.base_kernel, x1).to(device=dev, dtype=torch.float64))
            return out
        Q = E.alloc_matrix(x1.shape[0], x2.shape[0], dev)
        eng.gemm_nt(A1, A2, Q, alpha=1.0, beta=0.0)
        if self.training and same:
            kd = dense_cov_diag(self.base_kernel, x1).to(device=dev, dtype=torch.float64)
            corr = (kd - Q.diagonal()).clamp_min(0.0)
            Q.diagonal().add_(corr)
        return Q

    def forward(self, x1, x2, diag=False, **params):
        return self._forward64(x1, x2, diag=diag).to(device=x1.device, dtype=x1.dtype)


# gpytorch/kernels/kernel.py is a module in GPyTorch: `gpytorch.kernels.kernel` appears as a type annotation in the reference's
# tests/gp/test_spatiotemporal_gp.py:89
import sys as _sys

kernel = _sys.modules[__name__]
