"""gpytorch.settings stand-ins (only what BattGP touches).

``fast_pred_var`` / ``max_cholesky_size`` exist for API compatibility: this engine ALWAYS takes the exact Cholesky path
(what GPyTorch computes under ``max_cholesky_size(N+1)``), never CG/LOVE -- see DESIGN.md "Semantics vs GPyTorch defaults".
"""
from __future__ import annotations

import torch


class _Flag:
    _default = False
    _state = None

    def __init__(self, state: bool = True):
        self.state = bool(state)
        self.prev = None

    @classmethod
    def on(cls) -> bool:
        return cls._default if cls._state is None else cls._state

    @classmethod
    def off(cls) -> bool:
        return not cls.on()

    def __enter__(self):
        self.prev = type(self)._state
        type(self)._state = self.state
        return self

    def __exit__(self, *exc):
        type(self)._state = self.prev
        return False


class fast_pred_var(_Flag):
    """battcellgp_full.py:171, standard_models.py:40.  No effect: variances are exact either way."""
    _default = False
    _state = None


class debug(_Flag):
    """gpytorch.settings.debug(False) silences the train-input checks (test_standard_models.py:25)."""
    _default = True
    _state = None


class _Value:
    _default = None
    _value = None

    def __init__(self, value):
        self.v = value
        self.prev = None

    @classmethod
    def value(cls, *a):
        return cls._default if cls._value is None else cls._value

    def __enter__(self):
        self.prev = type(self)._value
        type(self)._value = self.v
        return self

    def __exit__(self, *exc):
        type(self)._value = self.prev
        return False


class max_cholesky_size(_Value):
    """Accepted and ignored: every size is factorised exactly."""
    _default = 800
    _value = None


class cholesky_max_tries(_Value):
    _default = 3
    _value = None


class min_variance(_Value):
    """Lower clamp of predictive variances: 1e-10 (fp64), 1e-6 (fp32), 1e-3 (fp16) as in GPyTorch."""
    _default = None
    _value = None

    @classmethod
    def value(cls, dtype=torch.float64):
        if cls._value is not None:
            return cls._value
        return {torch.float64: 1e-10, torch.float32: 1e-6, torch.float16: 1e-3}.get(dtype, 1e-10)
