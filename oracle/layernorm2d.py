"""CPU ORACLE (test infrastructure, NOT product code) -- LayerNorm2d (+ GELU).

Restates LayerNorm2d.forward (Models/modules.py:22-27): layer norm over the channel axis of a
(B, C, H, W) map (biased variance, eps inside the square root, affine weight/bias), and the
out_norm -> out_act(GELU, exact erf form) chain of SS2D (Models/vmamba.py:271,287), with the
analytic backward.  fp64 numpy.  Pinned to the reference class run on CPU
(tests/golden/layernorm2d.npz, made by tests/golden/make_golden_layernorm.py).
"""
import math

import numpy as np

try:                                    # scipy is in the image; fall back to math.erf otherwise
    from scipy.special import erf as _erf
except ImportError:                     # pragma: no cover
    _erf = np.vectorize(math.erf)


def gelu(z):
    return 0.5 * z * (1.0 + _erf(z / math.sqrt(2.0)))


def gelu_grad(z):
    return 0.5 * (1.0 + _erf(z / math.sqrt(2.0))) + z * np.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)


def layernorm2d_fwd(x, weight, bias=None, eps=1e-5, act="none"):
    """x (B, C, ...) -> y same shape; statistics over axis 1."""
    x = np.asarray(x, np.float64)
    shp = (1, -1) + (1,) * (x.ndim - 2)
    mu = x.mean(axis=1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=1, keepdims=True)
    z = (x - mu) / np.sqrt(var + eps) * np.asarray(weight, np.float64).reshape(shp)
    if bias is not None:
        z = z + np.asarray(bias, np.float64).reshape(shp)
    return gelu(z) if act == "gelu" else z


def layernorm2d_bwd(gy, x, weight, bias=None, eps=1e-5, act="none"):
    """-> (gx, gweight, gbias)."""
    x = np.asarray(x, np.float64)
    gy = np.asarray(gy, np.float64)
    shp = (1, -1) + (1,) * (x.ndim - 2)
    w = np.asarray(weight, np.float64).reshape(shp)
    mu = x.mean(axis=1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xh = (x - mu) * rstd
    if act == "gelu":
        z = xh * w + (0.0 if bias is None else np.asarray(bias, np.float64).reshape(shp))
        gy = gy * gelu_grad(z)
    red = tuple(i for i in range(x.ndim) if i != 1)
    gw = (gy * xh).sum(axis=red)
    gb = gy.sum(axis=red)
    g = gy * w
    gx = rstd * (g - g.mean(axis=1, keepdims=True) - xh * (g * xh).mean(axis=1, keepdims=True))
    return gx, gw, gb
