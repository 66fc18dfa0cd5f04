from ...engine import NanError, NotPSDError  # noqa: F401
