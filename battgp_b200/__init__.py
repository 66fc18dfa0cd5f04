"""battgp_b200 -- B200-native exact-GP engine for BattGP's ``full_gp`` path (fp64, sm_100a)."""
from . import _lib  # noqa: F401
from .engine import (Engine, FitState, KernelSpec, NanError, NotPSDError, NumericalWarning, Term, battgp_spec, fit,
                     get_engine, matern_periodic_spec, predict, scaled_rbf_spec)
from ._lib import MATERN52, PERIODIC, RBF, WIENER, BattGPLibraryError

__all__ = ["Engine", "FitState", "KernelSpec", "NanError", "NotPSDError", "NumericalWarning", "Term", "battgp_spec",
           "fit", "get_engine", "matern_periodic_spec", "predict", "scaled_rbf_spec", "MATERN52", "PERIODIC", "RBF",
           "WIENER", "BattGPLibraryError"]
