"""CPU oracle for BattGP's ``full_gp`` exact-GP path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  Nothing under ``battgp_b200/`` imports it; the product path fails loudly
when the CUDA library is missing.

It is a plain numpy/scipy (LAPACK ``dpotrf``/``dtrtrs``) fp64 restatement of what the reference computes
through GPyTorch when the Cholesky path is taken (``gpytorch.settings.max_cholesky_size(N+1)``):

* model / kernel structure ........ /root/reference/src/batt_models/cell_gp.py:27-36
  (zero mean, ``Scale(Wiener[dim 0]) + Scale(RBF-ARD[dims 1..3])``, Gaussian noise)
* Wiener covariance ............... /root/reference/src/gp/wiener_kernel.py:10-32
* isotropic Scale*RBF exact GP .... /root/reference/src/gp/standard_models.py:8-50
* predict semantics (latent f variance, no noise added; variance only)
                                    /root/reference/src/batt_models/battcellgp_full.py:168-195,
                                    /root/reference/src/gp/recursive_gp.py:120
* default hyper-parameters ........ /root/reference/src/config.py:39-43
* loss definition (-mll, mll = LML/N; callers scale by N)
                                    /root/reference/src/gp/training.py:27-43
* query grid ...................... /root/reference/src/batt_models/battgp_full.py:98,
                                    /root/reference/src/batt_models/battcellgp_full.py:199-206

GPyTorch itself (``gpytorch>=1.11``, /root/reference/requirements.txt:12 -- a floor, no lock file) is NOT
vendored in the reference and not installable here, so its arithmetic is restated from its published
formulas: RBF ``exp(-0.5*sqdist(x/l))`` with the mean-centred quadratic
expansion, ``ScaleKernel``, ``GaussianLikelihood`` noise on the train block, ``ExactMarginalLogLikelihood``
``= LML/N``, ``MultivariateNormal.variance`` clamp at 1e-10 (fp64).  Matern-5/2 and Periodic are NOT in
the reference at all (BASELINE.json config 3): their parity is UNPINNED and defined only by this file.

Pinning status (see tests/test_oracle.py and tests/golden/):
* posterior mean/variance: PINNED against the reference's analytic known-answer tests
  (/root/reference/tests/gp/test_standard_models.py:12-47) and against outputs of the reference's own
  ``RecursiveGP`` / ``ApproxSpatioTemporalGP`` code (/root/reference/src/gp/recursive_gp.py,
  spatiotemporal_gp.py) run in the build container -- the same cross-checks the reference's tests make
  (test_recursive_gp.py:195-232, test_spatiotemporal_gp.py:218-282); golden vectors in tests/golden/.
* ``WienerKernel.forward`` values: PINNED by executing the reference's wiener_kernel.py forward body
  (tests/golden/make_golden.py) on seeded inputs.
* LML value, LML gradient, L and alpha themselves: parity UNPINNED by any reference test (no reference
  test touches them); self-validated here by finite differences and the identities L L^T = K, K alpha = y.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np
import scipy.linalg as sla

# term types (shared numbering with include/battgp_b200.h)
WIENER, RBF, MATERN52, PERIODIC = 0, 1, 2, 3

MIN_VARIANCE_F64 = 1e-10  # gpytorch.settings.min_variance (fp64) [GPyTorch-recall, SURVEY Appendix C]


@dataclass
class Term:
    """One ``ScaleKernel(base)`` summand.  ``dims`` = GPyTorch ``active_dims``."""

    type: int
    dims: Sequence[int]
    outputscale: float
    lengthscale: Sequence[float] = ()   # per active dim (ARD); isotropic = repeat
    period: Sequence[float] = ()        # PERIODIC only


@dataclass
class KernelSpec:
    terms: list[Term] = field(default_factory=list)


def battgp_spec(outputscale_wiener=4.23e-13, outputscale_rbf=0.0099,
                lengthscale_rbf=(12.11, 33.75, 45.14)) -> KernelSpec:
    """cell_gp.py:32-36 with the config.py:39-43 defaults."""
    return KernelSpec([Term(WIENER, [0], outputscale_wiener),
                       Term(RBF, [1, 2, 3], outputscale_rbf, tuple(lengthscale_rbf))])


def scaled_rbf_spec(d: int, outputscale: float, lengthscale: float) -> KernelSpec:
    """standard_models.py:24 -- ScaleKernel(RBFKernel()) over all d columns, one lengthscale."""
    return KernelSpec([Term(RBF, list(range(d)), outputscale, (lengthscale,) * d)])


def matern_periodic_spec(s_m=0.0099, ls=(12.11, 33.75, 45.14), s_p=1e-4, period=1.0, ls_p=1.0) -> KernelSpec:
    """BASELINE.json config 3: Matern-5/2-ARD(I,SOC,T) + Periodic(t).  Not in reference."""
    return KernelSpec([Term(MATERN52, [1, 2, 3], s_m, tuple(ls)),
                       Term(PERIODIC, [0], s_p, (ls_p,), (period,))])


# --------------------------------------------------------------------------- covariance terms
def wiener_cov(t1: np.ndarray, t2: np.ndarray) -> np.ndarray:
    """wiener_kernel.py:32: min^3/3 + |t-t'| * min^2/2, min = minimum(t, t')."""
    t1 = np.asarray(t1, np.float64).reshape(-1, 1)
    t2 = np.asarray(t2, np.float64).reshape(1, -1)
    m = np.minimum(t1, t2)
    return m ** 3 / 3.0 + np.abs(t1 - t2) * m ** 2 / 2.0


def _scaled_sqdist(a: np.ndarray, b: np.ndarray, ls) -> np.ndarray:
    """sum_d ((a_d-b_d)/l_d)^2, difference form (what the CUDA kernel evaluates)."""
    ls = np.asarray(ls, np.float64)
    out = np.zeros((a.shape[0], b.shape[0]))
    for d in range(a.shape[1]):
        diff = (a[:, d:d + 1] - b[:, d:d + 1].T) / ls[d]
        out += diff * diff
    return out


def _scaled_sqdist_gpytorch(a: np.ndarray, b: np.ndarray, ls) -> np.ndarray:
    """[GPyTorch-recall] Kernel.covar_dist(square_dist=True): divide by l, subtract the column mean of a
    from both, one GEMM on [-2a, |a|^2, 1] x [b, 1, |b|^2]^T, zero the diagonal when a is b, clamp_min(0)."""
    ls = np.asarray(ls, np.float64)
    same = a is b or (a.shape == b.shape and np.array_equal(a, b))
    a_ = a / ls
    b_ = b / ls
    adj = a_.mean(axis=0, keepdims=True)
    a_ = a_ - adj
    b_ = b_ - adj
    an = (a_ ** 2).sum(1, keepdims=True)
    bn = (b_ ** 2).sum(1, keepdims=True)
    lhs = np.hstack([-2.0 * a_, an, np.ones_like(an)])
    rhs = np.hstack([b_, np.ones_like(bn), bn])
    res = lhs @ rhs.T
    if same:
        np.fill_diagonal(res, 0.0)
    return np.maximum(res, 0.0)


def term_cov(term: Term, x1: np.ndarray, x2: np.ndarray, *, gpytorch_expansion: bool = False) -> np.ndarray:
    a = np.asarray(x1, np.float64)[:, list(term.dims)]
    b = np.asarray(x2, np.float64)[:, list(term.dims)]
    if term.type == WIENER:
        return wiener_cov(a[:, 0], b[:, 0])
    if term.type == RBF:
        sq = (_scaled_sqdist_gpytorch if gpytorch_expansion else _scaled_sqdist)(a, b, term.lengthscale)
        return np.exp(-0.5 * sq)
    if term.type == MATERN52:
        # GPyTorch MaternKernel(nu=2.5): r = sqrt(clamp(sqdist, 1e-30)); (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r)
        r = np.sqrt(np.maximum(_scaled_sqdist(a, b, term.lengthscale), 1e-30))
        s5 = math.sqrt(5.0)
        return (1.0 + s5 * r + (5.0 / 3.0) * r * r) * np.exp(-s5 * r)
    if term.type == PERIODIC:
        # GPyTorch PeriodicKernel: exp(-2 sum_d sin^2(pi (a_d-b_d)/p_d) / l_d)  (divides by l, not l^2)
        acc = np.zeros((a.shape[0], b.shape[0]))
        for d in range(a.shape[1]):
            s = np.sin(math.pi * (a[:, d:d + 1] - b[:, d:d + 1].T) / term.period[d])
            acc += s * s / term.lengthscale[d]
        return np.exp(-2.0 * acc)
    raise ValueError(term.type)


def cov(spec: KernelSpec, x1: np.ndarray, x2: np.ndarray, **kw) -> np.ndarray:
    """k(x1, x2) = sum_i s_i * k_i(x1[:, dims_i], x2[:, dims_i])  -- ``kernel.forward(x1, x2)``."""
    out = np.zeros((x1.shape[0], x2.shape[0]))
    for t in spec.terms:
        out += t.outputscale * term_cov(t, x1, x2, **kw)
    return out


def cov_diag(spec: KernelSpec, x: np.ndarray) -> np.ndarray:
    """diag k(x, x): Wiener -> t^3/3; stationary terms -> 1."""
    x = np.asarray(x, np.float64)
    out = np.zeros(x.shape[0])
    for t in spec.terms:
        if t.type == WIENER:
            out += t.outputscale * x[:, t.dims[0]] ** 3 / 3.0
        else:
            out += t.outputscale
    return out


def train_cov(spec: KernelSpec, x: np.ndarray, noise: float) -> np.ndarray:
    """K = k(X, X) + sigma_n^2 I  (cell_gp.py:27 GaussianLikelihood)."""
    k = cov(spec, x, x)
    k[np.diag_indices_from(k)] += noise
    return k


# --------------------------------------------------------------------------- exact GP
@dataclass
class Fit:
    L: np.ndarray        # lower Cholesky factor of K
    alpha: np.ndarray    # K^-1 y
    z: np.ndarray        # L^-1 y
    lml: float           # log marginal likelihood (NOT divided by N)
    logdet: float        # log|K|
    jitter: float = 0.0


def cholesky_lower(k: np.ndarray, overwrite: bool = False) -> np.ndarray:
    """LAPACK dpotrf, lower.  Raises LinAlgError when not PD."""
    c, info = sla.lapack.dpotrf(k, lower=1, clean=1, overwrite_a=1 if overwrite else 0)
    if info != 0:
        raise np.linalg.LinAlgError(f"dpotrf info={info}")
    return c


def psd_safe_cholesky(k: np.ndarray, max_tries: int = 3):
    """[GPyTorch-recall] psd_safe_cholesky: retry with jitter 1e-8 * 10^i (fp64), i < max_tries."""
    try:
        return cholesky_lower(k), 0.0
    except np.linalg.LinAlgError:
        pass
    prev = 0.0
    kj = k.copy()
    for i in range(max_tries):
        jit = 1e-8 * 10 ** i
        kj[np.diag_indices_from(kj)] += jit - prev
        prev = jit
        try:
            return cholesky_lower(kj), jit
        except np.linalg.LinAlgError:
            continue
    raise np.linalg.LinAlgError("matrix not positive definite after jitter retries")


def fit(spec: KernelSpec, x: np.ndarray, y: np.ndarray, noise: float) -> Fit:
    y = np.asarray(y, np.float64).reshape(-1)
    n = y.shape[0]
    L, jit = psd_safe_cholesky(train_cov(spec, x, noise))
    z = sla.solve_triangular(L, y, lower=True)
    alpha = sla.solve_triangular(L, z, lower=True, trans="T")
    logdet = 2.0 * np.log(np.diag(L)).sum()
    lml = -0.5 * float(z @ z) - 0.5 * logdet - 0.5 * n * math.log(2.0 * math.pi)
    return Fit(L, alpha, z, lml, logdet, jit)


def predict(spec: KernelSpec, x: np.ndarray, f: Fit, xq: np.ndarray, full_cov: bool = False,
            clamp: bool = True):
    """Latent-f posterior: mean = K*N alpha; var = diag k** - colsum((L^-1 K_N*)^2); no noise added."""
    kq = cov(spec, xq, x)                       # M x N
    mean = kq @ f.alpha
    v = sla.solve_triangular(f.L, kq.T, lower=True)   # N x M
    if full_cov:
        c = cov(spec, xq, xq) - v.T @ v
        return mean, c
    var = cov_diag(spec, xq) - (v * v).sum(axis=0)
    if clamp:
        var = np.maximum(var, MIN_VARIANCE_F64)
    return mean, var


def mll_loss(f: Fit, n: int) -> float:
    """training.py:40: loss = -mll, mll = LML / N (ExactMarginalLogLikelihood)."""
    return -f.lml / n


# --------------------------------------------------------------------------- gradient (K9)
def lml_grad(spec: KernelSpec, x: np.ndarray, y: np.ndarray, noise: float) -> dict:
    """dLML/dtheta = 0.5 tr((alpha alpha^T - K^-1) dK/dtheta) for noise, every outputscale and every
    lengthscale / period.  Returns {"noise": g, "terms": [{"outputscale": g,
    "lengthscale": [...], "period": [...]}, ...], "lml": value}."""
    x = np.asarray(x, np.float64)
    f = fit(spec, x, y, noise)
    n = x.shape[0]
    kinv = sla.lapack.dpotri(f.L, lower=1)[0]
    kinv = np.tril(kinv) + np.tril(kinv, -1).T
    q = 0.5 * (np.outer(f.alpha, f.alpha) - kinv)
    out = {"noise": float(np.trace(q)), "terms": [], "lml": f.lml}
    for t in spec.terms:
        a = x[:, list(t.dims)]
        kt = term_cov(t, x, x)
        g = {"outputscale": float((q * kt).sum()), "lengthscale": [], "period": []}
        for j in range(len(t.lengthscale)):
            diff = a[:, j:j + 1] - a[:, j:j + 1].T
            l = t.lengthscale[j]
            if t.type == RBF:
                dk = kt * diff ** 2 / l ** 3
            elif t.type == MATERN52:
                r = np.sqrt(np.maximum(_scaled_sqdist(a, a, t.lengthscale), 1e-30))
                s5 = math.sqrt(5.0)
                # dk/dr = -(5/3) r (1 + sqrt5 r) exp(-sqrt5 r); dr/dl_j = -diff_j^2 / (l_j^3 r)
                dk = (5.0 / 3.0) * (1.0 + s5 * r) * np.exp(-s5 * r) * diff ** 2 / l ** 3
            elif t.type == PERIODIC:
                p = t.period[j]
                s = np.sin(math.pi * diff / p)
                dk = kt * 2.0 * s * s / l ** 2
                dkp = kt * (4.0 / l) * s * np.cos(math.pi * diff / p) * math.pi * diff / p ** 2
                g["period"].append(float(t.outputscale * (q * dkp).sum()))
            else:
                continue
            g["lengthscale"].append(float(t.outputscale * (q * dk).sum()))
        out["terms"].append(g)
    return out


# --------------------------------------------------------------------------- synthetic data
def synth_field_data(n: int, seed: int = 0):
    """Synthetic 8s1p field telemetry, statistics mirrored from the reference fixtures."""
    rng = np.random.default_rng(seed)
    t = np.sort(np.round(rng.uniform(0, 120, n) * 17280) / 17280)
    cur = rng.uniform(-80, -5, n)
    soc = rng.uniform(40, 95, n)
    temp = np.clip(rng.normal(24, 4, n), 10, 35)
    x = np.ascontiguousarray(np.stack([t, cur, soc, temp], axis=1))
    y = (4e-3 * (1 + 2e-3 * t) + 1e-3 * np.exp(-(temp - 10) / 15) + 5e-4 * (soc - 70) ** 2 / 900
         - 1e-5 * cur / 80 + rng.normal(0, math.sqrt(2.33e-6), n))
    return x, y


def query_grid(x: np.ndarray, m: int = 300, op=(-15.0, 90.0, 25.0)) -> np.ndarray:
    """battgp_full.py:98 + battcellgp_full.py:199-206; op = gp_runner.py:32."""
    t = np.linspace(x[0, 0], x[-1, 0], m)
    return np.column_stack([t, np.full(m, op[0]), np.full(m, op[1]), np.full(m, op[2])])
