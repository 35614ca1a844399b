"""fsw_gnn_b200 - B200-native Fourier Sliced-Wasserstein embedding / message passing.

Drop-in for the hot path of tal-amir/fsw-gnn: `FSW_embedding`, `FSW_conv`, `FSW_readout`,
`segcumsum` keep the reference's signatures and run on libfsw_embedding.so (sm_100a CUDA, C ABI in
include/fsw_embedding.h).  No CPU fallback: the library must be built (`python -m fsw_gnn_b200.build`).
"""
from . import _lib  # noqa: F401
from .fsw_embedding import FSW_embedding, minimize_mutual_coherence, segcumsum, segcumsum_slow, sp  # noqa: F401
from .fsw_conv import FSW_conv, FSW_readout  # noqa: F401

__all__ = ["FSW_embedding", "FSW_conv", "FSW_readout", "segcumsum", "segcumsum_slow", "minimize_mutual_coherence", "sp"]
