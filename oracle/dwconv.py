"""CPU ORACLE (test infrastructure, NOT product code) -- depthwise conv (+ bias, + SiLU).

Restates `self.act(self.conv2d(x))` of SS2D.forward (Models/vmamba.py:281-285) for
nn.Conv2d(C, C, k, padding=(k-1)//2, groups=C, bias=conv_bias) (vmamba.py:72-79) with act = SiLU,
plus the analytic backward.  fp64 numpy.  Pinned to torch's own conv2d/SiLU (the reference's ops)
through tests/golden/dwconv.npz (made by tests/golden/make_golden_dwconv.py).
"""
import numpy as np


def _silu(z):
    return z / (1.0 + np.exp(-z))


def _silu_grad(z):
    s = 1.0 / (1.0 + np.exp(-z))
    return s * (1.0 + z * (1.0 - s))


def _conv(x, w):
    """x (B, C, H, W), w (C, k, k) -> same-size cross-correlation with zero padding."""
    B, C, H, W = x.shape
    k = w.shape[-1]
    p = k // 2
    xp = np.zeros((B, C, H + 2 * p, W + 2 * p), np.float64)
    xp[:, :, p:p + H, p:p + W] = x
    y = np.zeros((B, C, H, W), np.float64)
    for i in range(k):
        for j in range(k):
            y += w[None, :, i, j, None, None] * xp[:, :, i:i + H, j:j + W]
    return y


def dwconv_fwd(x, weight, bias=None, act="none"):
    x = np.asarray(x, np.float64)
    w = np.asarray(weight, np.float64).reshape(x.shape[1], weight.shape[-1], weight.shape[-1])
    z = _conv(x, w)
    if bias is not None:
        z = z + np.asarray(bias, np.float64)[None, :, None, None]
    return _silu(z) if act == "silu" else z


def dwconv_bwd(dy, x, weight, bias=None, act="none"):
    """-> (dx, dweight (C,1,k,k), dbias (C))."""
    x = np.asarray(x, np.float64)
    g = np.asarray(dy, np.float64)
    C = x.shape[1]
    k = weight.shape[-1]
    p = k // 2
    w = np.asarray(weight, np.float64).reshape(C, k, k)
    if act == "silu":
        z = _conv(x, w) + (0.0 if bias is None else np.asarray(bias, np.float64)[None, :, None, None])
        g = g * _silu_grad(z)
    dx = _conv(g, w[:, ::-1, ::-1])
    B, _, H, W = x.shape
    xp = np.zeros((B, C, H + 2 * p, W + 2 * p), np.float64)
    xp[:, :, p:p + H, p:p + W] = x
    dw = np.zeros((C, k, k), np.float64)
    for i in range(k):
        for j in range(k):
            dw[:, i, j] = (g * xp[:, :, i:i + H, j:j + W]).sum(axis=(0, 2, 3))
    return dx, dw.reshape(C, 1, k, k), g.sum(axis=(0, 2, 3))
