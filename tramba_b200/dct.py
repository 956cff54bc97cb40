"""DCT2D high/low-frequency split -- same module surface as Models/DCT_2D.py:6-77.

`DCT2D(height, width).forward(x) -> (high, low)`; the bases live in registered buffers
`dct_x.weight` (W x W) and `dct_y.weight` (H x H), so reference checkpoints load unchanged
(state_dict keys).  The transform itself runs in libtramba_b200.so.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib


def _basis(n):
    if isinstance(_lib.lib, _lib._NotLoaded):
        # bench.py --impl reference (library deliberately not mapped): the buffer is a host constant,
        # W[v, j] = cos(pi (j + 1/2) v / n) / sqrt(n) * (sqrt(2) if v > 0)   (DCT_2D.py:37-45)
        v, j = np.arange(n, dtype=np.float64).reshape(n, 1), np.arange(n, dtype=np.float64).reshape(1, n)
        w = np.cos(np.pi * (j + 0.5) * v / n) / np.sqrt(n) * np.where(v > 0, np.sqrt(2.0), 1.0)
        return torch.from_numpy(w.astype(np.float32))
    w = np.empty((n, n), dtype=np.float32)
    _lib.lib.tramba_dct_basis(n, w.ctypes.data)  # DCT_2D.py:37-45, bit-exact
    return torch.from_numpy(w)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _DCT2DSplit(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, wy, wx):
        ctx.in_dtype = x.dtype
        if not x.is_cuda:
            raise RuntimeError("tramba_b200 has no CPU path: DCT2D needs a CUDA tensor")
        B, C, H, W = x.shape
        # the reference multiplies x by fp32 basis buffers, so any input is promoted to fp32
        x = x.contiguous().float()
        wy, wx = wy.contiguous().float(), wx.contiguous().float()
        high = x.new_empty((B, C, H // 2, W // 2))
        low = x.new_empty((B, C, H // 2, W // 2))
        with torch.cuda.device(x.device), _lib.timed("dct2d_fwd", x.numel() * 6):
            _lib.check(_lib.lib.tramba_dct2d_fwd(x.data_ptr(), wy.data_ptr(), wx.data_ptr(), high.data_ptr(),
                                                 low.data_ptr(), B * C, H, W, _stream()))
        ctx.save_for_backward(wy, wx)
        ctx.hw = (H, W)
        return high, low

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, ghigh, glow):
        wy, wx = ctx.saved_tensors
        H, W = ctx.hw
        B, C = ghigh.shape[:2]
        ghigh, glow = ghigh.contiguous().float(), glow.contiguous().float()
        gx = ghigh.new_empty((B, C, H, W))
        with torch.cuda.device(ghigh.device), _lib.timed("dct2d_bwd", gx.numel() * 6):
            _lib.check(_lib.lib.tramba_dct2d_bwd(ghigh.data_ptr(), glow.data_ptr(), wy.data_ptr(), wx.data_ptr(),
                                                 gx.data_ptr(), B * C, H, W, _stream()))
        return gx.to(ctx.in_dtype), None, None


def dct2d_split(x, wy, wx):
    return _DCT2DSplit.apply(x, wy, wx)


class DCT2DSpatialTransformLayer_x(nn.Module):
    """Holds the width basis (DCT_2D.py:32-45); the contraction is done by DCT2D.forward."""

    def __init__(self, width):
        super().__init__()
        self.register_buffer("weight", _basis(width))


class DCT2DSpatialTransformLayer_y(nn.Module):
    """Holds the height basis (DCT_2D.py:56-69)."""

    def __init__(self, height):
        super().__init__()
        self.register_buffer("weight", _basis(height))


class DCT2D(nn.Module):
    def __init__(self, height, width):
        super().__init__()
        self.dct_x = DCT2DSpatialTransformLayer_x(width)
        self.dct_y = DCT2DSpatialTransformLayer_y(height)

    def forward(self, x):
        if x.shape[2] % 2 or x.shape[3] % 2:
            raise ValueError("tramba_b200.DCT2D: H and W must be even for the quadrant split")
        H, W = x.shape[2], x.shape[3]
        if x.is_cuda and not _lib.lib.tramba_dct2d_supported(H, W):
            raise RuntimeError(f"tramba_b200.DCT2D: no kernel for {H} x {W} planes (tramba_dct2d_supported); there is "
                               f"no library fallback")
        return dct2d_split(x, self.dct_y.weight, self.dct_x.weight)  # (high, low), DCT_2D.py:12-18
