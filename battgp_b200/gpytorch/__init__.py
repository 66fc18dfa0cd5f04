"""Drop-in for the part of ``gpytorch`` that BattGP's full_gp path uses, executing on the
battgp_b200 CUDA engine.  Activate with ``import battgp_b200.shim; battgp_b200.shim.install()`` or by putting
``<repo>/shims`` on PYTHONPATH; /root/reference/src then runs unmodified."""
from . import constraints, distributions, kernels, likelihoods, means, mlls, models, settings, utils  # noqa: F401
from .module import Module  # noqa: F401
from .mlls import ExactMarginalLogLikelihood  # noqa: F401

__version__ = "1.11+battgp_b200"
