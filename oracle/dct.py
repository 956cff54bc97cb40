"""CPU ORACLE (test infrastructure, NOT product code) -- DCT2D high/low split.

Restates Models/DCT_2D.py: orthonormal DCT-II bases (lines 37-45 for x, 61-69 for y; built
in Python doubles, stored as fp32 buffers), y = dct_y(dct_x(x)) (12-16), and the quadrant
split low = Y[:H/2,:W/2], high = Y[H/2:,W/2:], returned as (high, low) (20-29).
Pinned to the reference module run on CPU (tests/golden/dct.npz) and to scipy.fft.dctn.
"""
import math
import numpy as np


def dct_basis(n):
    """W[v, j] = cos(pi*(0.5+j)*v/n)/sqrt(n) * (sqrt(2) if v else 1), rounded to fp32 as the
    reference's registered buffer is (DCT_2D.py:38-45)."""
    w = np.zeros((n, n), np.float32)
    for v in range(n):
        for j in range(n):
            b = math.cos(math.pi * (0.5 + j) * v / n) / math.sqrt(n)
            if v != 0:
                b = b * math.sqrt(2)
            w[v, j] = b
    return w


def dct2d_full(x, dtype=np.float64):
    """x: (B, C, H, W) -> Y = Wy . X . Wx^T per plane."""
    H, W = x.shape[-2:]
    wy = dct_basis(H).astype(dtype)
    wx = dct_basis(W).astype(dtype)
    return np.einsum("uh,bchw,vw->bcuv", wy, x.astype(dtype), wx, optimize=True)


def dct2d_split(x, dtype=np.float64):
    """Returns (high, low) like DCT2D.forward (DCT_2D.py:12-18)."""
    y = dct2d_full(x, dtype)
    hh, hw = y.shape[2] // 2, y.shape[3] // 2
    return y[:, :, hh:, hw:], y[:, :, :hh, :hw]


def dct2d_split_bwd(ghigh, glow, H, W, dtype=np.float64):
    """Adjoint of dct2d_split: dX = Wy^T . dY . Wx with dY zero outside the two quadrants."""
    B, C = ghigh.shape[:2]
    hh, hw = H // 2, W // 2
    gy = np.zeros((B, C, H, W), dtype)
    gy[:, :, hh:, hw:] = ghigh
    gy[:, :, :hh, :hw] = glow
    wy = dct_basis(H).astype(dtype)
    wx = dct_basis(W).astype(dtype)
    return np.einsum("uh,bcuv,vw->bchw", wy, gy, wx, optimize=True)
