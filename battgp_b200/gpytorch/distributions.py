"""gpytorch.distributions.MultivariateNormal (cell_gp.py:62; read at battcellgp_full.py:175,180, standard_models.py:43-48)."""
from __future__ import annotations

import warnings

import torch

from . import settings
from ..engine import NumericalWarning


class MultivariateNormal:
    """mean + (possibly lazy) covariance.  ``covar`` may be a dense tensor, a kernels.LazyKernelMatrix (prior) or a
    models.PosteriorCovariance (posterior); all expose ``to_dense()`` and ``diagonal()``."""

    def __init__(self, mean, covariance_matrix, validate_args=False):
        self.loc = mean
        self._covar = covariance_matrix

    @property
    def mean(self):
        return self.loc

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        return self._covar if torch.is_tensor(self._covar) else self._covar.to_dense()

    @property
    def variance(self):
        var = self._covar.diagonal(dim1=-1, dim2=-2) if torch.is_tensor(self._covar) else self._covar.diagonal()
        min_variance = settings.min_variance.value(var.dtype)
        if bool((var < min_variance).any()):
            warnings.warn(f"Negative variance values detected. This is likely due to numerical instabilities. "
                          f"Rounding negative variances up to {min_variance}.", NumericalWarning)
            var = var.clamp_min(min_variance)
        return var

    @property
    def stddev(self):
        return self.variance.sqrt()

    def confidence_region(self):
        s2 = self.stddev * 2.0
        return self.mean - s2, self.mean + s2

    @property
    def event_shape(self):
        return self.loc.shape[-1:]

    @property
    def batch_shape(self):
        return self.loc.shape[:-1]

    def __repr__(self):
        return f"MultivariateNormal(loc: {tuple(self.loc.shape)})"
