from .core import FactorizedTensor  # noqa: F401
