"""gpytorch.utils.warnings (gp_runner.py:14,28 filters NumericalWarning)."""
from ...engine import NumericalWarning  # noqa: F401


class GPInputWarning(UserWarning):
    """Raised when a model in eval mode is called on its own training inputs (test_standard_models.py:24-25)."""


class OldVersionWarning(UserWarning):
    pass
