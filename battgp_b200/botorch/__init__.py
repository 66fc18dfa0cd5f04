"""Drop-in for the two botorch names /root/reference/src/gp/training.py uses (:1,6,8,82-95)."""
from . import fit, settings  # noqa: F401
