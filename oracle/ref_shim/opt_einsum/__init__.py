"""Stand-in for `opt_einsum`: only imported by the reference's complex-half einsum helper
(`neuralop/layers/einsum_utils.py`), which the full-precision path never calls."""


def contract_path(*args, **kwargs):
    raise NotImplementedError("opt_einsum shim: complex-half path is out of scope for the oracle")
