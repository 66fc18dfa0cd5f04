"""gpytorch.means (cell_gp.py:30, standard_models.py:23, recursive_gp.py:45,53,131)."""
from __future__ import annotations

import torch

from .module import Module


class Mean(Module):
    def forward(self, x):
        raise NotImplementedError

    def __call__(self, x):
        if x.dim() == 1:
            x = x.unsqueeze(-1)
        return self.forward(x)


class ZeroMean(Mean):
    def __init__(self, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.batch_shape = batch_shape

    def forward(self, x):
        return torch.zeros(x.shape[:-1], dtype=x.dtype, device=x.device)


class ConstantMean(Mean):
    def __init__(self, constant_prior=None, constant_constraint=None, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.register_parameter("raw_constant", torch.nn.Parameter(torch.zeros(batch_shape)))
        if constant_constraint is not None:
            self.register_constraint("raw_constant", constant_constraint)

    @property
    def constant(self):
        return self._get_constrained("raw_constant")

    @constant.setter
    def constant(self, value):
        self._set_constrained("raw_constant", value)

    def forward(self, x):
        return self.constant.to(x.dtype).expand(x.shape[:-1])
