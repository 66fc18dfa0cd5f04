"""Tensor-level host side of the B200 exact-GP engine: torch owns memory and streams, every FLOP of the path runs in
libbattgp_b200.so through the C-ABI (include/battgp_b200.h).  No CPU fallback, nothing imported from ``oracle/``.

Mirrors what GPyTorch executes for BattGP's ``full_gp`` mode under the Cholesky path:
  fit      = ExactGP prediction-strategy set-up / ExactMarginalLogLikelihood
             (/root/reference/src/batt_models/battcellgp_full.py:173, /root/reference/src/gp/training.py:40)
  predict  = MultivariateNormal.mean / .variance of the latent f (battcellgp_full.py:175-180)
"""
from __future__ import annotations

import ctypes as C
import math
import os
import warnings
from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import MATERN52, PERIODIC, RBF, WIENER, BgpKernelSpec

MIN_VARIANCE_F64 = 1e-10      # gpytorch.settings.min_variance for fp64
JITTERS_F64 = (1e-8, 1e-7, 1e-6)  # psd_safe_cholesky retries


class NotPSDError(RuntimeError):
    """Cholesky failed even after the jitter retries (GPyTorch: linear_operator.utils.errors.NotPSDError)."""


class NanError(RuntimeError):
    """NaN encountered in the covariance matrix (GPyTorch: linear_operator.utils.errors.NanError)."""


class NumericalWarning(RuntimeWarning):
    """gpytorch.utils.warnings.NumericalWarning (silenced by /root/reference/gp_runner.py:28)."""


@dataclass
class Term:
    """One ``ScaleKernel(base)`` summand; ``dims`` = GPyTorch ``active_dims``."""
    type: int
    dims: Sequence[int]
    outputscale: float
    lengthscale: Sequence[float] = ()
    period: Sequence[float] = ()


@dataclass
class KernelSpec:
    terms: list = field(default_factory=list)

    def to_c(self, noise: float = 0.0) -> BgpKernelSpec:
        if not 1 <= len(self.terms) <= _lib.MAX_TERMS:
            raise ValueError(f"1..{_lib.MAX_TERMS} kernel terms supported, got {len(self.terms)}")
        s = BgpKernelSpec()
        s.nterms = len(self.terms)
        s.noise = float(noise)
        for i, t in enumerate(self.terms):
            ct = s.terms[i]
            ct.type = int(t.type)
            dims = list(t.dims)
            if len(dims) > _lib.MAX_DIMS or len(dims) < 1:
                raise ValueError("1..8 active dims per term")
            ct.ndims = len(dims)
            for k, d in enumerate(dims):
                ct.dims[k] = int(d)
            ct.outputscale = float(t.outputscale)
            ls = list(t.lengthscale)
            if t.type != WIENER:
                if len(ls) == 1 and len(dims) > 1:
                    ls = ls * len(dims)
                if len(ls) != len(dims):
                    raise ValueError("lengthscale must have one entry per active dim")
            for k, v in enumerate(ls[: _lib.MAX_DIMS]):
                ct.lengthscale[k] = float(v)
            per = list(t.period)
            if t.type == PERIODIC:
                if len(per) == 1 and len(dims) > 1:
                    per = per * len(dims)
                if len(per) != len(dims):
                    raise ValueError("period must have one entry per active dim")
            for k, v in enumerate(per[: _lib.MAX_DIMS]):
                ct.period[k] = float(v)
        return s


def battgp_spec(outputscale_wiener=4.23e-13, outputscale_rbf=0.0099, lengthscale_rbf=(12.11, 33.75, 45.14)) -> KernelSpec:
    """cell_gp.py:32-36 with the config.py:39-43 defaults."""
    return KernelSpec([Term(WIENER, [0], outputscale_wiener), Term(RBF, [1, 2, 3], outputscale_rbf, tuple(lengthscale_rbf))])


def scaled_rbf_spec(d: int, outputscale: float, lengthscale: float) -> KernelSpec:
    """standard_models.py:24."""
    return KernelSpec([Term(RBF, list(range(d)), outputscale, (lengthscale,) * d)])


def matern_periodic_spec(s_m=0.0099, ls=(12.11, 33.75, 45.14), s_p=1e-4, period=1.0, ls_p=1.0) -> KernelSpec:
    """BASELINE.json config 3."""
    return KernelSpec([Term(MATERN52, [1, 2, 3], s_m, tuple(ls)), Term(PERIODIC, [0], s_p, (ls_p,), (period,))])


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _check_f64_cuda(t: torch.Tensor, name: str, device: torch.device):
    if t.dtype != torch.float64:
        raise TypeError(f"{name}: expected float64, got {t.dtype}")
    if t.device != device:
        raise ValueError(f"{name}: expected device {device}, got {t.device}")


def alloc_matrix(rows: int, cols: int, device) -> torch.Tensor:
    """rows x cols fp64 view with an even leading dimension (16-byte aligned rows for cp.async / vector stores)."""
    ld = cols + (cols & 1)
    return torch.empty((rows, ld), dtype=torch.float64, device=device)[:, :cols]


class Engine:
    """One per CUDA device (one process per GPU in the reference's deployment, gp_runner.py:150-171)."""

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.BattGPLibraryError("battgp_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.L = _lib.lib()
        torch.cuda.init()
        h = C.c_void_p()
        _lib.check(self.L.bgp_ctx_create(self.device.index, C.byref(h)), "bgp_ctx_create")
        self.h = h
        # fp64-accurate int8/tcgen05 trailing updates (csrc/ozaki.cu) are on by default; BATTGP_OZAKI=0 keeps every
        # contraction on the FP64 DMMA pipe
        self.ozaki = False
        self.set("ozaki", 0 if os.environ.get("BATTGP_OZAKI", "1") == "0" else 1)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.bgp_ctx_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set(self, key: str, value: int):
        _lib.check(self.L.bgp_ctx_set(self.h, key.encode(), int(value)), f"bgp_ctx_set({key})")
        if key == "ozaki":
            self.ozaki = bool(value)

    def _ensure_workspace(self, n: int):
        """Scratch for the int8/tcgen05 trailing updates: a torch tensor (so empty_cache() can release it), re-used."""
        if not getattr(self, "ozaki", False):
            return
        need = int(self.L.bgp_potrf_workspace_bytes(self.h, n))
        ws = getattr(self, "_ws", None)
        if need and (ws is None or ws.numel() < need):
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            _lib.check(self.L.bgp_ctx_set_workspace(self.h, _ptr(self._ws), self._ws.numel()), "bgp_ctx_set_workspace")

    def ensure_workspace_bytes(self, need: int):
        """Make at least ``need`` bytes of int8-path scratch available (no-op when the path is off)."""
        if not getattr(self, "ozaki", False) or need <= 0:
            return
        ws = getattr(self, "_ws", None)
        if ws is None or ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=self.device)
            _lib.check(self.L.bgp_ctx_set_workspace(self.h, _ptr(self._ws), self._ws.numel()), "bgp_ctx_set_workspace")

    def release_workspace(self):
        _lib.check(self.L.bgp_ctx_set_workspace(self.h, C.c_void_p(0), 0), "bgp_ctx_set_workspace")
        self._ws = None

    def kernel_profile(self, enable: bool):
        """Timed CUDA events around the trailing-update launches of bgp_potrf (bench.py's roofline).  Turning it off
        returns {"ms", "flop", "launches"} summed over the launches recorded since it was turned on."""
        if enable:
            _lib.check(self.L.bgp_ctx_kernel_profile(self.h, 1), "bgp_ctx_kernel_profile")
            return None
        ms, fl, nl = C.c_double(0.0), C.c_double(0.0), C.c_int64(0)
        _lib.check(self.L.bgp_ctx_kernel_profile_read(self.h, C.byref(ms), C.byref(fl), C.byref(nl)), "bgp_ctx_kernel_profile_read")
        _lib.check(self.L.bgp_ctx_kernel_profile(self.h, 0), "bgp_ctx_kernel_profile")
        return {"ms": ms.value, "flop": fl.value, "launches": int(nl.value)}

    @property
    def launches(self) -> int:
        return int(self.L.bgp_ctx_launches(self.h))

    @staticmethod
    def _ld(t: torch.Tensor) -> int:
        if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
            raise ValueError("expected a row-major 2-D tensor")
        return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])

    # ------------------------------------------------------------------ K1-K3
    def cov_build(self, spec: KernelSpec, x1: torch.Tensor, x2: Optional[torch.Tensor] = None, *, noise: float = 0.0,
                  symmetric: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _check_f64_cuda(x1, "x1", self.device)
        x1 = x1.contiguous() if x1.stride(-1) != 1 else x1
        if symmetric:
            x2 = x1
        else:
            if x2 is None:
                raise ValueError("x2 required unless symmetric")
            _check_f64_cuda(x2, "x2", self.device)
            x2 = x2.contiguous() if x2.stride(-1) != 1 else x2
        n1, n2 = x1.shape[0], x2.shape[0]
        if out is None:
            out = alloc_matrix(n1, n2, self.device)
        cs = spec.to_c(noise)
        rc = self.L.bgp_cov_build(self.h, C.byref(cs), _ptr(x1), n1, self._ld(x1), _ptr(x2), n2, self._ld(x2),
                                  _ptr(out), self._ld(out), 1 if symmetric else 0, self._stream())
        _lib.check(rc, "bgp_cov_build")
        return out

    def cov_diag(self, spec: KernelSpec, x: torch.Tensor) -> torch.Tensor:
        _check_f64_cuda(x, "x", self.device)
        x = x.contiguous() if x.stride(-1) != 1 else x
        out = torch.empty(x.shape[0], dtype=torch.float64, device=self.device)
        cs = spec.to_c(0.0)
        _lib.check(self.L.bgp_cov_diag(self.h, C.byref(cs), _ptr(x), x.shape[0], self._ld(x), _ptr(out), self._stream()),
                   "bgp_cov_diag")
        return out

    # ------------------------------------------------------------------ dense block
    def gemm_nt(self, A: torch.Tensor, B: torch.Tensor, C_: torch.Tensor, alpha=1.0, beta=0.0, tri=False, roff=0, coff=0):
        M, K = A.shape
        N = B.shape[0]
        assert B.shape[1] == K and C_.shape == (M, N)
        rc = self.L.bgp_gemm_nt(self.h, M, N, K, float(alpha), _ptr(A), self._ld(A), _ptr(B), self._ld(B), float(beta),
                                _ptr(C_), self._ld(C_), 1 if tri else 0, roff, coff, self._stream())
        _lib.check(rc, "bgp_gemm_nt")
        return C_

    def gemm_nt_i8(self, A: torch.Tensor, B: torch.Tensor, C_: torch.Tensor, alpha=1.0, tri=False, roff=0, coff=0, work=None):
        """EXPERIMENTAL: C += alpha * A B^T through int8 slicing on tcgen05 (Ozaki scheme)."""
        M, K = A.shape
        N = B.shape[0]
        need = int(self.L.bgp_gemm_nt_i8_work_bytes(M, N, K))
        if work is None or work.numel() < need:
            work = torch.empty(need, dtype=torch.uint8, device=self.device)
        rc = self.L.bgp_gemm_nt_i8(self.h, M, N, K, float(alpha), _ptr(A), self._ld(A), _ptr(B), self._ld(B), _ptr(C_),
                                   self._ld(C_), 1 if tri else 0, roff, coff, _ptr(work), work.numel(), self._stream())
        _lib.check(rc, "bgp_gemm_nt_i8")
        return C_

    def oz_slice(self, P: torch.Tensor, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        rows, K = P.shape
        need = int(self.L.bgp_oz_slice_bytes(rows, K))
        if buf is None or buf.numel() < need:
            buf = torch.empty(need, dtype=torch.uint8, device=self.device)
        _lib.check(self.L.bgp_oz_slice(self.h, _ptr(P), rows, K, self._ld(P), _ptr(buf), buf.numel(), self._stream()), "bgp_oz_slice")
        return buf

    def oz_slice_gather(self, P: torch.Tensor, rows: int, blkmap: torch.Tensor, blkrows: int, buf: Optional[torch.Tensor] = None):
        """Slice ``rows`` logical rows whose block b lives at source block blkmap[b] of P (int32 device tensor)."""
        K = P.shape[1]
        need = int(self.L.bgp_oz_slice_bytes(rows, K))
        if buf is None or buf.numel() < need:
            buf = torch.empty(need, dtype=torch.uint8, device=self.device)
        _lib.check(self.L.bgp_oz_slice_gather(self.h, _ptr(P), rows, K, self._ld(P), _ptr(blkmap), blkrows, _ptr(buf), buf.numel(),
                                              self._stream()), "bgp_oz_slice_gather")
        return buf

    def oz_gemm(self, bufA, rowsA, arow0, bufB, rowsB, brow0, C_, K, alpha=-1.0, tri=False, roff=0, coff=0, tpc=0):
        """``tpc`` tiles per CTA (0 = one persistent CTA per SM; > 0 lets CTAs retire so that a concurrent high-priority
        stream finds free SMs -- the sharded look-ahead)."""
        M, N = C_.shape
        if tpc != getattr(self, "_oz_tpc_gemm", 0):
            self.set("oz_tpc_gemm", tpc)
            self._oz_tpc_gemm = tpc
        rc = self.L.bgp_oz_gemm(self.h, _ptr(bufA), rowsA, arow0, _ptr(bufB), rowsB, brow0, M, N, K, float(alpha), _ptr(C_),
                                self._ld(C_), 1 if tri else 0, roff, coff, self._stream())
        _lib.check(rc, "bgp_oz_gemm")
        return C_

    # components of the modular (CRT) int8 emulation (csrc/ozaki2.cu; groundwork, not used by potrf yet)
    def oz2_residues(self, A: torch.Tensor):
        """A [rows, K] fp64 -> (residues [16, rows, K] int8, exponents [rows] int32)."""
        _check_f64_cuda(A, "A", self.device)
        rows, K = A.shape
        res = torch.empty((16, rows, K), dtype=torch.int8, device=self.device)
        expo = torch.empty(rows, dtype=torch.int32, device=self.device)
        rc = self.L.bgp_oz2_residues(self.h, _ptr(A), rows, K, self._ld(A), _ptr(res), _ptr(expo), self._stream())
        _lib.check(rc, "bgp_oz2_residues")
        return res, expo

    def oz2_crt(self, G: torch.Tensor, ea: torch.Tensor, eb: torch.Tensor, C_: torch.Tensor, alpha: float = 1.0):
        """G [16, M, N] int32 (contiguous) -> C += alpha * 2^(ea_i + eb_j - 110) * CRT(G)."""
        _check_f64_cuda(C_, "C", self.device)
        if G.dtype != torch.int32 or not G.is_contiguous() or G.shape[0] != 16 or tuple(G.shape[1:]) != tuple(C_.shape):
            raise ValueError("oz2_crt: G must be a contiguous int32 [16, M, N] tensor matching C")
        M, N = C_.shape
        rc = self.L.bgp_oz2_crt(self.h, _ptr(G), M, N, _ptr(ea), _ptr(eb), float(alpha), _ptr(C_), self._ld(C_), self._stream())
        _lib.check(rc, "bgp_oz2_crt")
        return C_

    def oz2_gemm(self, A: torch.Tensor, B: torch.Tensor, C_: torch.Tensor, alpha: float = 1.0):
        """EXPERIMENTAL: C += alpha * A B^T through the modular scheme end to end (csrc/next/ozaki2_mma.cu)."""
        for t, nm in ((A, "A"), (B, "B"), (C_, "C")):
            _check_f64_cuda(t, nm, self.device)
        M, K = A.shape
        N = B.shape[0]
        need = int(self.L.bgp_oz2_gemm_work_bytes(M, N, K))
        work = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        off = (-work.data_ptr()) % 256
        rc = self.L.bgp_oz2_gemm(self.h, M, N, K, float(alpha), _ptr(A), self._ld(A), _ptr(B), self._ld(B), _ptr(C_), self._ld(C_),
                                 C.c_void_p(work.data_ptr() + off), need, self._stream())
        _lib.check(rc, "bgp_oz2_gemm")
        return C_

    # ------------------------------------------------------------------ K4
    def potrf(self, A: torch.Tensor):
        """In-place lower Cholesky.  ``A`` is n x n, or (n + mx) x n with mx extra right-hand-side rows that leave as
        X L^-T (bgp_potrf_aug).  Returns (info, logdet, dinv)."""
        _check_f64_cuda(A, "A", self.device)
        n = A.shape[1]
        mx = A.shape[0] - n
        if mx < 0:
            raise ValueError("potrf: expected n x n or (n + mx) x n")
        dinv = torch.empty(int(self.L.bgp_potrf_dinv_elems(n)), dtype=torch.float64, device=self.device)
        self._ensure_workspace(n + mx)
        logdet = C.c_double(0.0)
        rc = self.L.bgp_potrf_aug(self.h, _ptr(A), n, mx, self._ld(A), _ptr(dinv), C.byref(logdet), self._stream())
        _lib.check(rc, "bgp_potrf_aug")
        return rc, logdet.value, dinv

    def potrf_async(self, A: torch.Tensor, info_dev: torch.Tensor, logdet_dev: torch.Tensor) -> torch.Tensor:
        """Asynchronous in-place Cholesky of an n x n or (n + mx) x n matrix (small systems; no host sync, results stay on
        the device): returns the 128-block inverses."""
        n = A.shape[1]
        mx = A.shape[0] - n
        dinv = torch.empty(int(self.L.bgp_potrf_dinv_elems(n)), dtype=torch.float64, device=self.device)
        rc = self.L.bgp_potrf_async(self.h, _ptr(A), n, mx, self._ld(A), _ptr(dinv), _ptr(info_dev), _ptr(logdet_dev), self._stream())
        _lib.check(rc, "bgp_potrf_async")
        return dinv

    def lml_dev(self, z: torch.Tensor, logdet_dev: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        _lib.check(self.L.bgp_lml_dev(self.h, _ptr(z), z.shape[0], _ptr(logdet_dev), _ptr(out), self._stream()), "bgp_lml_dev")
        return out

    # ------------------------------------------------------------------ K5
    def potrs_vec(self, Lm: torch.Tensor, dinv: torch.Tensor, y: torch.Tensor):
        n = Lm.shape[0]
        _check_f64_cuda(y, "y", self.device)
        y = y.contiguous()
        z = torch.empty(n, dtype=torch.float64, device=self.device)
        alpha = torch.empty(n, dtype=torch.float64, device=self.device)
        rc = self.L.bgp_potrs_vec(self.h, _ptr(Lm), n, self._ld(Lm), _ptr(dinv), _ptr(y), _ptr(z), _ptr(alpha),
                                  self._stream())
        _lib.check(rc, "bgp_potrs_vec")
        return z, alpha

    # ------------------------------------------------------------------ K7
    def trsm_rlt(self, Lm: torch.Tensor, dinv: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
        """X <- X L^-T in place (rows of X are the right-hand sides)."""
        n = Lm.shape[0]
        assert X.shape[1] == n
        rc = self.L.bgp_trsm_rlt(self.h, _ptr(Lm), n, self._ld(Lm), _ptr(dinv), _ptr(X), X.shape[0], self._ld(X),
                                 self._stream())
        _lib.check(rc, "bgp_trsm_rlt")
        return X

    def predict_tail(self, Kq=None, alpha=None, V=None, kdiag=None, min_var=MIN_VARIANCE_F64, mean_out=None, var_out=None):
        ref = Kq if Kq is not None else V
        m, n = ref.shape
        mean = (mean_out if mean_out is not None else torch.empty(m, dtype=torch.float64, device=self.device)) if Kq is not None else None
        var = (var_out if var_out is not None else torch.empty(m, dtype=torch.float64, device=self.device)) if V is not None else None
        rc = self.L.bgp_predict_tail(self.h, m, n, _ptr(Kq), self._ld(Kq) if Kq is not None else 0, _ptr(alpha),
                                     _ptr(V), self._ld(V) if V is not None else 0, _ptr(kdiag), float(min_var),
                                     _ptr(mean), _ptr(var), self._stream())
        _lib.check(rc, "bgp_predict_tail")
        return mean, var

    # ------------------------------------------------------------------ sharded building blocks
    def potrf_block(self, A: torch.Tensor, info_dev: torch.Tensor, logdet_dev: torch.Tensor) -> torch.Tensor:
        """Asynchronous in-place Cholesky of one diagonal block; returns its 128-block inverses."""
        nb = A.shape[0]
        dinv = torch.empty(int(self.L.bgp_potrf_dinv_elems(nb)), dtype=torch.float64, device=self.device)
        rc = self.L.bgp_potrf_block(self.h, _ptr(A), nb, self._ld(A), _ptr(dinv), _ptr(info_dev), _ptr(logdet_dev), self._stream())
        _lib.check(rc, "bgp_potrf_block")
        return dinv

    def trsv(self, Lm: torch.Tensor, dinv: torch.Tensor, b: torch.Tensor, trans: bool = False) -> torch.Tensor:
        _lib.check(self.L.bgp_trsv(self.h, _ptr(Lm), Lm.shape[0], self._ld(Lm), _ptr(dinv), _ptr(b), 1 if trans else 0,
                                   self._stream()), "bgp_trsv")
        return b

    def gemv_t(self, A: torch.Tensor, v: torch.Tensor, y: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
        _lib.check(self.L.bgp_gemv_t(self.h, _ptr(A), A.shape[0], A.shape[1], self._ld(A), _ptr(v), _ptr(y), float(alpha),
                                     self._stream()), "bgp_gemv_t")
        return y

    def rowsumsq(self, V: torch.Tensor, out: torch.Tensor, accumulate: bool = False) -> torch.Tensor:
        _lib.check(self.L.bgp_rowsumsq(self.h, _ptr(V), V.shape[0], V.shape[1], self._ld(V), _ptr(out), 1 if accumulate else 0,
                                       self._stream()), "bgp_rowsumsq")
        return out

    # ------------------------------------------------------------------ K8
    def lml(self, z: torch.Tensor, logdet: float) -> float:
        out = C.c_double(0.0)
        _lib.check(self.L.bgp_lml(self.h, _ptr(z), z.shape[0], float(logdet), C.byref(out), self._stream()), "bgp_lml")
        return out.value

    # ------------------------------------------------------------------ K9
    def potri(self, Lm: torch.Tensor, dinv: torch.Tensor, work: Optional[torch.Tensor] = None) -> torch.Tensor:
        """L (lower, in place) -> K^-1 (lower triangle of the same storage)."""
        n = Lm.shape[0]
        if work is None:
            work = alloc_matrix(n, n, self.device)
        rc = self.L.bgp_potri(self.h, _ptr(Lm), n, self._ld(Lm), _ptr(dinv), _ptr(work), self._ld(work), self._stream())
        _lib.check(rc, "bgp_potri")
        return Lm

    def lml_grad(self, spec: KernelSpec, noise: float, x: torch.Tensor, Kinv: torch.Tensor, alpha: torch.Tensor):
        cs = spec.to_c(noise)
        slots = _lib.check(self.L.bgp_grad_slots(C.byref(cs)), "bgp_grad_slots")
        g = torch.empty(slots, dtype=torch.float64, device=self.device)
        x = x.contiguous() if x.stride(-1) != 1 else x
        rc = self.L.bgp_lml_grad(self.h, C.byref(cs), _ptr(x), x.shape[0], self._ld(x), _ptr(Kinv), self._ld(Kinv),
                                 _ptr(alpha), _ptr(g), self._stream())
        _lib.check(rc, "bgp_lml_grad")
        return g


_engines: dict = {}


def get_engine(device) -> Engine:
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.BattGPLibraryError(
            f"battgp_b200 needs a CUDA device (got {device}); there is no CPU fallback for the exact-GP path")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _engines:
        _engines[idx] = Engine(torch.device("cuda", idx))
    return _engines[idx]


# ---------------------------------------------------------------------------------------------------------------
@dataclass
class FitState:
    """What GPyTorch's DefaultPredictionStrategy caches: L (as the lower triangle of the K buffer) and alpha."""
    spec: KernelSpec
    noise: float
    x: torch.Tensor
    L: torch.Tensor
    dinv: torch.Tensor
    alpha: torch.Tensor
    z: torch.Tensor
    logdet: float
    lml: float
    jitter: float = 0.0
    xq: Optional[torch.Tensor] = None       # query points whose solve V = K_*N L^-T came out of the factorisation
    V: Optional[torch.Tensor] = None


def fit(spec: KernelSpec, x: torch.Tensor, y: torch.Tensor, noise: float, *, K_out: Optional[torch.Tensor] = None,
        kbuilder=None, potrf_events=None, xq: Optional[torch.Tensor] = None, kcross=None) -> FitState:
    """build K -> Cholesky (with GPyTorch's jitter retries) -> alpha -> LML.  ``kbuilder(out, extra_noise)`` may
    replace the fused build for kernels the engine does not know (it must fill the lower triangle of ``out``).

    With ``xq`` (M query points) the cross-covariance rows K_*N are appended under K and ride through the factorisation
    (bgp_potrf_aug): the state then already holds V = K_*N L^-T for the predictive variance -- BattGP builds a model,
    predicts once at 300 points and frees it (battgp_full.py:98-120), so fit and predict are one pass here."""
    eng = get_engine(x.device)
    n = x.shape[0]
    m = 0 if xq is None else xq.shape[0]
    Kfull = K_out if K_out is not None else alloc_matrix(n + m, n, x.device)
    if Kfull.shape[0] != n + m:
        raise ValueError(f"K_out must have {n + m} rows")
    K = Kfull[:n]
    jitter = 0.0
    for attempt in range(len(JITTERS_F64) + 1):
        if kbuilder is None:
            eng.cov_build(spec, x, noise=noise + jitter, symmetric=True, out=K)
        else:
            kbuilder(K, noise + jitter)
        if m:
            if kcross is None:
                eng.cov_build(spec, xq, x, out=Kfull[n:])
            else:
                Kfull[n:].copy_(kcross)
        if potrf_events is not None:
            potrf_events[0].record()
        info, logdet, dinv = eng.potrf(Kfull)
        if potrf_events is not None:
            potrf_events[1].record()
        if info == 0 and math.isfinite(logdet):
            break
        if info == 0 and not math.isfinite(logdet):
            raise NanError("cholesky: NaN/inf encountered in the covariance matrix")
        # a NaN pivot is reported as "not positive" by the leaf kernel: GPyTorch's psd_safe_cholesky raises NanError at once
        # instead of burning three more N^3 factorisations on jitter retries (the factor of a failed attempt is garbage, so
        # the inputs are what gets checked: X, the noise and one covariance entry for the hyper-parameters)
        if attempt == 0 and kbuilder is None:
            bad = not bool(torch.isfinite(x).all()) or not math.isfinite(noise)
            if not bad:
                probe = eng.cov_build(spec, x[:1], x[:1])
                bad = not bool(torch.isfinite(probe).all())
            if bad:
                raise NanError("cholesky: NaN/inf encountered in the covariance matrix")
        if attempt == len(JITTERS_F64):
            raise NotPSDError(f"Matrix not positive definite after repeatedly adding jitter up to {jitter:.1e} "
                              f"(first failing pivot {info}).")
        jitter = JITTERS_F64[attempt]
        warnings.warn(f"A not p.d., added jitter of {jitter:.1e} to the diagonal", NumericalWarning)
    z, alpha = eng.potrs_vec(K, dinv, y)
    lml = eng.lml(z, logdet)
    if not math.isfinite(lml):
        raise NanError("cholesky: NaN/inf encountered in the covariance matrix or the targets")
    st = FitState(spec, noise, x, K, dinv, alpha, z, logdet, lml, jitter)
    if m:
        st.xq, st.V = xq, Kfull[n:]
    return st


def predict(st: FitState, xq: torch.Tensor, *, full_cov: bool = False, clamp: bool = True, kcross=None, kdiag=None):
    """Latent-f posterior mean and variance (no noise added): battcellgp_full.py:168-195, recursive_gp.py:120."""
    eng = get_engine(st.x.device)
    Kq = eng.cov_build(st.spec, xq, st.x) if kcross is None else kcross
    mean, _ = eng.predict_tail(Kq=Kq, alpha=st.alpha)
    if st.V is not None and st.xq is not None and (st.xq is xq or (st.xq.shape == xq.shape and torch.equal(st.xq, xq))):
        V = st.V                                 # solved during the factorisation (fit(..., xq=...))
    else:
        V = eng.trsm_rlt(st.L, st.dinv, Kq)      # in place: Kq now holds K_*N L^-T
    if full_cov:
        m = xq.shape[0]
        Cq = eng.cov_build(st.spec, xq, xq) if kdiag is None else kdiag
        eng.gemm_nt(V, V, Cq, alpha=-1.0, beta=1.0)
        return mean, Cq
    kd = eng.cov_diag(st.spec, xq) if kdiag is None else kdiag
    _, var = eng.predict_tail(V=V, kdiag=kd, min_var=MIN_VARIANCE_F64 if clamp else -math.inf)
    return mean, var


def residual(st: FitState, y: torch.Tensor, block: int = 2048) -> float:
    """Matrix-free check at any size: ||K alpha - y|| / ||y|| with K REBUILT from X row-block by row-block (full width),
    so it does not depend on the factor that produced alpha."""
    eng = get_engine(st.x.device)
    n = st.x.shape[0]
    a_row = alloc_matrix(1, n, st.x.device)
    a_row[0].copy_(st.alpha)
    rr = torch.zeros(1, dtype=torch.float64, device=st.x.device)
    for b0 in range(0, n, block):
        e = min(n, b0 + block)
        krow = eng.cov_build(st.spec, st.x[b0:e], st.x)
        r = alloc_matrix(e - b0, 1, st.x.device)
        r[:, 0].copy_((st.noise + st.jitter) * st.alpha[b0:e] - y[b0:e])
        eng.gemm_nt(krow, a_row, r, alpha=1.0, beta=1.0)
        rt = alloc_matrix(1, e - b0, st.x.device)
        rt[0].copy_(r[:, 0])
        eng.rowsumsq(rt, rr, True)
    return math.sqrt(float(rr.item())) / float(torch.linalg.vector_norm(y).item())
