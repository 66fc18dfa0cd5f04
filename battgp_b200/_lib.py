"""ctypes binding of libbattgp_b200.so (include/battgp_b200.h).

The product path has NO fallback: if the library is missing or a call fails, this module raises.  Nothing here
imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(_HERE, "lib", "libbattgp_b200.so")
HEADER = os.path.join(_HERE, "..", "include", "battgp_b200.h")

WIENER, RBF, MATERN52, PERIODIC = 0, 1, 2, 3
MAX_TERMS, MAX_DIMS = 4, 8
E_ARG, E_CUDA, E_SPEC = -1, -2, -3


class BgpTerm(C.Structure):
    _fields_ = [("type", C.c_int32), ("ndims", C.c_int32), ("dims", C.c_int32 * MAX_DIMS),
                ("outputscale", C.c_double), ("lengthscale", C.c_double * MAX_DIMS),
                ("period", C.c_double * MAX_DIMS)]


class BgpKernelSpec(C.Structure):
    _fields_ = [("nterms", C.c_int32), ("_pad", C.c_int32), ("terms", BgpTerm * MAX_TERMS), ("noise", C.c_double)]


class BattGPLibraryError(RuntimeError):
    """libbattgp_b200.so missing / call failed.  There is deliberately no CPU fallback."""


_P = C.c_void_p
_I64 = C.c_int64
_D = C.c_double
_SPEC = C.POINTER(BgpKernelSpec)

# name -> (restype, argtypes); must list every function declared in include/battgp_b200.h
SIGNATURES = {
    "bgp_version": (C.c_int, []),
    "bgp_last_error": (C.c_char_p, []),
    "bgp_grad_slots": (C.c_int, [_SPEC]),
    "bgp_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "bgp_ctx_destroy": (None, [_P]),
    "bgp_ctx_set": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "bgp_potrf_workspace_bytes": (_I64, [_P, _I64]),
    "bgp_ctx_set_workspace": (C.c_int, [_P, _P, _I64]),
    "bgp_ctx_launches": (_I64, [_P]),
    "bgp_ctx_kernel_profile": (C.c_int, [_P, C.c_int]),
    "bgp_ctx_kernel_profile_read": (C.c_int, [_P, C.POINTER(_D), C.POINTER(_D), C.POINTER(_I64)]),
    "bgp_cov_build": (C.c_int, [_P, _SPEC, _P, _I64, _I64, _P, _I64, _I64, _P, _I64, C.c_int, _P]),
    "bgp_cov_diag": (C.c_int, [_P, _SPEC, _P, _I64, _I64, _P, _P]),
    "bgp_gemm_nt": (C.c_int, [_P, _I64, _I64, _I64, _D, _P, _I64, _P, _I64, _D, _P, _I64, C.c_int, _I64, _I64, _P]),
    "bgp_oz_slice_bytes": (_I64, [_I64, _I64]),
    "bgp_oz_slice": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I64, _P]),
    "bgp_oz_slice_gather": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I64, _P, _I64, _P]),
    "bgp_oz_gemm": (C.c_int, [_P, _P, _I64, _I64, _P, _I64, _I64, _I64, _I64, _I64, _D, _P, _I64, C.c_int, _I64, _I64, _P]),
    "bgp_panel_schedule": (C.c_int, [_I64, _I64, C.c_int, C.c_int, C.POINTER(_I64), C.c_int]),
    "bgp_oz2_residues": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _P]),
    "bgp_oz2_crt": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _D, _P, _I64, _P]),
    "bgp_oz2_gemm_work_bytes": (_I64, [_I64, _I64, _I64]),
    "bgp_oz2_gemm": (C.c_int, [_P, _I64, _I64, _I64, _D, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P]),
    "bgp_gemm_nt_i8_work_bytes": (_I64, [_I64, _I64, _I64]),
    "bgp_gemm_nt_i8": (C.c_int, [_P, _I64, _I64, _I64, _D, _P, _I64, _P, _I64, _P, _I64, C.c_int, _I64, _I64, _P, _I64, _P]),
    "bgp_potrf_dinv_elems": (_I64, [_I64]),
    "bgp_potrf": (C.c_int, [_P, _P, _I64, _I64, _P, C.POINTER(_D), _P]),
    "bgp_potrf_aug": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, C.POINTER(_D), _P]),
    "bgp_potrf_async": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _P, _P]),
    "bgp_lml_dev": (C.c_int, [_P, _P, _I64, _P, _P, _P]),
    "bgp_potrf_block": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _P, _P]),
    "bgp_potrs_vec": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _P, _P, _P]),
    "bgp_trsm_rlt": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _I64, _I64, _P]),
    "bgp_predict_tail": (C.c_int, [_P, _I64, _I64, _P, _I64, _P, _P, _I64, _P, _D, _P, _P, _P]),
    "bgp_lml": (C.c_int, [_P, _P, _I64, _D, C.POINTER(_D), _P]),
    "bgp_trsv": (C.c_int, [_P, _P, _I64, _I64, _P, _P, C.c_int, _P]),
    "bgp_gemv_t": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _D, _P]),
    "bgp_rowsumsq": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, C.c_int, _P]),
    "bgp_fault_eval": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _D, _D, _P, _P, _P, _P, _P, _P, _P, _P]),
    "bgp_potri": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _I64, _P]),
    "bgp_lml_grad": (C.c_int, [_P, _SPEC, _P, _I64, _I64, _P, _I64, _P, _P, _P]),
}


def declared_symbols(header: str = HEADER) -> list[str]:
    """Function names declared in the public header (used by the CPU test that checks the exports)."""
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bgp_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise BattGPLibraryError(
                f"{LIBPATH} not found: build it with `python -m battgp_b200.build` (needs nvcc). "
                "battgp_b200 has no CPU fallback.")
        try:
            L = C.CDLL(LIBPATH)
        except OSError as e:  # e.g. libcudart missing
            raise BattGPLibraryError(f"cannot load {LIBPATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str) -> int:
    """Raise on negative status; pass through 0 / positive (LAPACK-style info)."""
    if rc < 0:
        msg = {E_ARG: "bad argument", E_SPEC: "malformed kernel spec"}.get(rc)
        if rc == E_CUDA:
            msg = "CUDA error: " + (lib().bgp_last_error() or b"").decode()
        raise BattGPLibraryError(f"{what} failed ({rc}): {msg}")
    return rc
