"""Make ``import gpytorch`` / ``import botorch`` resolve to the battgp_b200 stand-ins."""
from __future__ import annotations

import importlib
import sys

_GP_SUB = ["constraints", "distributions", "kernels", "likelihoods", "means", "mlls", "models", "settings", "utils",
           "utils.warnings", "utils.errors", "module"]
_BO_SUB = ["fit", "settings"]


def install(force: bool = False) -> None:
    """Register the stand-ins in sys.modules.  A real GPyTorch already imported is left alone unless ``force``."""
    if "gpytorch" in sys.modules and not force and not getattr(sys.modules["gpytorch"], "__version__", "").endswith("battgp_b200"):
        raise RuntimeError("a real gpytorch is already imported; call install(force=True) to shadow it")
    g = importlib.import_module("battgp_b200.gpytorch")
    sys.modules["gpytorch"] = g
    for s in _GP_SUB:
        sys.modules["gpytorch." + s] = importlib.import_module("battgp_b200.gpytorch." + s)
    b = importlib.import_module("battgp_b200.botorch")
    sys.modules["botorch"] = b
    for s in _BO_SUB:
        sys.modules["botorch." + s] = importlib.import_module("battgp_b200.botorch." + s)
