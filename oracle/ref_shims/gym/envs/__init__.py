from .registration import make, register  # noqa: F401
