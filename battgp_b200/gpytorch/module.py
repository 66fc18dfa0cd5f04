"""Base module: raw parameters + constraints, GPyTorch style (``x`` <-> ``raw_x`` through ``raw_x_constraint``)."""
from __future__ import annotations

import torch


class Module(torch.nn.Module):
    def register_constraint(self, param_name: str, constraint):
        setattr(self, param_name + "_constraint", constraint)

    def _get_constrained(self, raw_name: str) -> torch.Tensor:
        raw = getattr(self, raw_name)
        c = getattr(self, raw_name + "_constraint", None)
        return c.transform(raw) if c is not None else raw

    def _set_constrained(self, raw_name: str, value) -> None:
        raw = getattr(self, raw_name)
        if not torch.is_tensor(value):
            value = torch.as_tensor(value)
        value = value.to(dtype=raw.dtype, device=raw.device)
        c = getattr(self, raw_name + "_constraint", None)
        if c is not None:
            if not c.check(value):
                raise RuntimeError(
                    f"Attempting to manually set a parameter value that is out of bounds of its current constraints, {c}. "
                    "Most likely, you want to do the following:\n likelihood = GaussianLikelihood"
                    "(noise_constraint=gpytorch.constraints.GreaterThan(better_lower_bound))")
            value = c.inverse_transform(value)
        with torch.no_grad():
            if value.numel() == raw.numel():
                raw.copy_(value.reshape(raw.shape))
            else:
                raw.copy_(value.expand_as(raw))

    def initialize(self, **kwargs):
        for name, val in kwargs.items():
            if name.startswith("raw_") and hasattr(self, name):
                p = getattr(self, name)
                with torch.no_grad():
                    v = torch.as_tensor(val).to(p)
                    p.copy_(v.reshape(p.shape) if v.numel() == p.numel() else v.expand_as(p))
            elif hasattr(type(self), name) and isinstance(getattr(type(self), name), property):
                setattr(self, name, val)
            elif "." in name:
                mod, _, rest = name.partition(".")
                getattr(self, mod).initialize(**{rest: val})
            else:
                raise AttributeError(f"Unknown parameter {name} for {type(self).__name__}")
        return self

    def hyperparameters(self):
        return self.parameters()

    def named_hyperparameters(self):
        return self.named_parameters()
