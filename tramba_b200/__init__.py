"""tramba_b200 -- B200-native (sm_100a) implementation of Tramba's 2D selective-scan hot path:
cross-scan -> selective-scan recurrence (fwd/bwd) -> cross-merge, plus the DCT2D high/low split,
behind the reference's own operator / module interface.

Importing the package loads libtramba_b200.so (built from tramba_b200/csrc with nvcc); there is
no CPU or PyTorch fallback for the compute entry points.
"""
import sys

from . import _lib  # noqa: F401  (fails loudly if the CUDA library cannot be loaded)
from . import selective_scan_cuda_oflex
from .dct import DCT2D, dct2d_split
from .dwconv import dwconv2d
from .norm import layer_norm2d
from .ops import (CrossMerge, CrossMerge_Ab_1direction, CrossMerge_Ab_2direction, CrossMerge_Diagonal,
                  CrossMerge_Dilation, CrossMerge_DS, CrossMerge_Hilbert, CrossMerge_Line,
                  CrossMerge_Line_4direction, CrossMerge_Spiral, CrossMerge_Window, CrossScan,
                  CrossScan_Ab_1direction, CrossScan_Ab_2direction, CrossScan_Diagonal, CrossScan_Dilation,
                  CrossScan_DS, CrossScan_Hilbert, CrossScan_Line, CrossScan_Line_4direction, CrossScan_Spiral,
                  CrossScan_Window, SelectiveScanOflex, cross_merge, cross_scan, make_custom_pair,
                  selective_scan_fn)
from .ss2d import SS2D, LayerNorm2d, Linear2d


def install_extension():
    """Register our extension module under the bare name the reference imports
    (`import selective_scan_cuda_oflex`, Models/SS2D/csms6s.py:741-745)."""
    sys.modules["selective_scan_cuda_oflex"] = selective_scan_cuda_oflex
    return selective_scan_cuda_oflex


__all__ = [
    "SS2D", "DCT2D", "Linear2d", "LayerNorm2d", "SelectiveScanOflex", "selective_scan_fn",
    "CrossScan", "CrossMerge", "CrossScan_Dilation", "CrossMerge_Dilation", "CrossScan_Window",
    "CrossMerge_Window", "CrossScan_Line", "CrossMerge_Line", "cross_scan", "cross_merge", "dct2d_split",
    "CrossScan_Line_4direction", "CrossMerge_Line_4direction", "CrossScan_Spiral", "CrossMerge_Spiral",
    "CrossScan_Hilbert", "CrossMerge_Hilbert", "CrossScan_Diagonal", "CrossMerge_Diagonal", "CrossScan_DS",
    "CrossMerge_DS", "CrossScan_Ab_2direction", "CrossMerge_Ab_2direction", "CrossScan_Ab_1direction",
    "CrossMerge_Ab_1direction", "make_custom_pair",
    "layer_norm2d", "dwconv2d",
    "install_extension", "selective_scan_cuda_oflex",
]
