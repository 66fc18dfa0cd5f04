"""gpytorch.constraints: Interval / Positive / GreaterThan / LessThan with GPyTorch's transforms
(sigmoid for two-sided intervals, softplus for one-sided) -- /root/reference/src/gpytorch_utils.py:17-80 builds these."""
from __future__ import annotations

import math

import torch
from torch.nn.functional import softplus


def inv_softplus(x: torch.Tensor) -> torch.Tensor:
    return x + torch.log(-torch.expm1(-x))


def inv_sigmoid(x: torch.Tensor) -> torch.Tensor:
    return torch.log(x) - torch.log1p(-x)


class Interval(torch.nn.Module):
    def __init__(self, lower_bound, upper_bound, transform=torch.sigmoid, inv_transform=inv_sigmoid, initial_value=None):
        super().__init__()
        lower_bound = torch.as_tensor(lower_bound).to(torch.get_default_dtype())
        upper_bound = torch.as_tensor(upper_bound).to(torch.get_default_dtype())
        if torch.any(lower_bound >= upper_bound):
            raise ValueError("Got parameter bounds with empty intervals.")
        self.register_buffer("lower_bound", lower_bound)
        self.register_buffer("upper_bound", upper_bound)
        self._transform = transform
        self._inv_transform = inv_transform
        self.initial_value = initial_value

    @property
    def enforced(self) -> bool:
        return self._transform is not None

    def check(self, tensor) -> bool:
        return bool(torch.all(tensor <= self.upper_bound.to(tensor)) and torch.all(tensor >= self.lower_bound.to(tensor)))

    def check_raw(self, tensor) -> bool:
        return self.check(self.transform(tensor))

    def _two_sided(self):
        return bool(torch.all(torch.isfinite(self.lower_bound)) and torch.all(torch.isfinite(self.upper_bound)))

    def transform(self, tensor: torch.Tensor) -> torch.Tensor:
        if not self.enforced or not self._two_sided():
            return tensor
        lo, hi = self.lower_bound.to(tensor), self.upper_bound.to(tensor)
        return self._transform(tensor) * (hi - lo) + lo

    def inverse_transform(self, tensor: torch.Tensor) -> torch.Tensor:
        if not self.enforced or not self._two_sided():
            return tensor
        lo, hi = self.lower_bound.to(tensor), self.upper_bound.to(tensor)
        return self._inv_transform((tensor - lo) / (hi - lo))

    def __repr__(self):
        return f"{type(self).__name__}({self.lower_bound.tolist()!r}, {self.upper_bound.tolist()!r})"


class GreaterThan(Interval):
    def __init__(self, lower_bound, transform=softplus, inv_transform=inv_softplus, initial_value=None):
        super().__init__(lower_bound, math.inf, transform, inv_transform, initial_value)

    def transform(self, tensor):
        return self._transform(tensor) + self.lower_bound.to(tensor) if self.enforced else tensor

    def inverse_transform(self, tensor):
        return self._inv_transform(tensor - self.lower_bound.to(tensor)) if self.enforced else tensor


class Positive(GreaterThan):
    def __init__(self, transform=softplus, inv_transform=inv_softplus, initial_value=None):
        super().__init__(0.0, transform, inv_transform, initial_value)

    def transform(self, tensor):
        return self._transform(tensor) if self.enforced else tensor

    def inverse_transform(self, tensor):
        return self._inv_transform(tensor) if self.enforced else tensor


class LessThan(Interval):
    def __init__(self, upper_bound, transform=softplus, inv_transform=inv_softplus, initial_value=None):
        super().__init__(-math.inf, upper_bound, transform, inv_transform, initial_value)

    def transform(self, tensor):
        return -self._transform(-tensor) + self.upper_bound.to(tensor) if self.enforced else tensor

    def inverse_transform(self, tensor):
        return -self._inv_transform(-(tensor - self.upper_bound.to(tensor))) if self.enforced else tensor
