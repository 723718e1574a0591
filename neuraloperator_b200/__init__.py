"""neuraloperator_b200: B200-native SpectralConv forward+backward (sm_100a kernels behind a C ABI).

Only the spectral-convolution hot path of neuraloperator lives here; use `SpectralConv` as the
`conv_module=` of the reference's FNO / TFNO / FNOBlocks.
"""
from .spectral_conv import (BaseSpectralConv, Plan, SpectralConv, analyze, contract_dense,  # noqa: F401
                            contract_dense_backward, get_plan, spectral_conv_dense, synthesize)
from .factorized import FactorizedWeight  # noqa: F401
from .data_parallel import GradientAllReducer, PeerGradientAllReducer  # noqa: F401
from .integration import use_b200_layers  # noqa: F401

__version__ = "0.1.0"
from .fno_block import (ChannelMLP, ComplexValued, Flattened1dConv, FNOBlocks, SoftGating, channel_mix,  # noqa: F401
                        set_tensor_core_mixing, skip_connection, uses_tensor_core_mixing)
