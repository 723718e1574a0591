from .factorized_tensors.core import FactorizedTensor  # noqa: F401
