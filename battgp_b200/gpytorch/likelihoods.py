"""gpytorch.likelihoods.GaussianLikelihood (cell_gp.py:27,64-81; standard_models.py:17,26; recursive_gp.py:82)."""
from __future__ import annotations

import torch

from .constraints import GreaterThan
from .distributions import MultivariateNormal
from .module import Module


class HomoskedasticNoise(Module):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size()):
        super().__init__()
        if noise_constraint is None:
            noise_constraint = GreaterThan(1e-4)      # GPyTorch's default; BattGP replaces it (battcellgp_full.py:71)
        self.register_parameter("raw_noise", torch.nn.Parameter(torch.zeros(*batch_shape, 1)))
        self.register_constraint("raw_noise", noise_constraint)

    @property
    def noise(self):
        return self._get_constrained("raw_noise")

    @noise.setter
    def noise(self, value):
        self._set_constrained("raw_noise", value)

    def __float__(self):
        return float(self.noise.detach().reshape(-1)[0])


class Likelihood(Module):
    pass


class GaussianLikelihood(Likelihood):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.noise_covar = HomoskedasticNoise(noise_prior, noise_constraint, batch_shape)

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    @property
    def raw_noise(self):
        return self.noise_covar.raw_noise

    def forward(self, function_samples, *params, **kwargs):
        raise NotImplementedError("sampling through the likelihood is outside the exact-GP hot path")

    def __call__(self, function_dist, *params, **kwargs):
        """marginal: p(y) = N(mean, K + sigma_n^2 I)"""
        if isinstance(function_dist, MultivariateNormal):
            cov = function_dist._covar
            noise = self.noise
            if torch.is_tensor(cov):
                cov = cov + torch.diag_embed(noise.to(cov).expand(cov.shape[-1]))
            else:
                cov = cov.add_diagonal(noise)
            return MultivariateNormal(function_dist.mean, cov)
        return super().__call__(function_dist, *params, **kwargs)
