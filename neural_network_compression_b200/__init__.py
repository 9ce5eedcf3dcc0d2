"""B200-native prune + 1-D k-means weight-compression hot path.

Drop-in for the compression helpers of angelocatalani/neural-network-compression
(`neural_network_compression.common.utility`): hand-written sm_100a CUDA kernels behind the C ABI declared in
include/nnc.h, loaded through ctypes.  `from neural_network_compression_b200.common import utility`.
"""
from . import _native  # noqa: F401
from .common import utility  # noqa: F401

__version__ = "0.1.0"
