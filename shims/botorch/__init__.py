"""PYTHONPATH shim: ``import botorch`` -> battgp_b200.botorch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
del sys.modules[__name__]
import battgp_b200.shim as _s  # noqa: E402

_s.install(force=True)
