"""CPU ORACLE (test infrastructure, NOT product code) -- selective scan fwd/bwd.

ctypes front-end over oracle/sscan_ref.c plus a tiny pure-numpy statement used to
cross-check the C code.  PARITY UNPINNED against the reference's own (absent) extension; pinned instead to an
independent executed implementation of the same published algorithm (HF transformers MambaMixer.slow_forward,
tests/test_oracle_sscan_pin.py) -- see the sscan_ref.c header.  Follows the call sites
Models/SS2D/csms6s.py:904-923 and the argument preparation in Models/vmamba.py:244-257.
"""
import ctypes
import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _prep(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def selective_scan_fwd(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=True,
                       dtype=np.float64):
    """u, delta: (b, dim, L); A: (dim, N); B, C: (b, G, N, L); D, delta_bias: (dim).
    Returns out (b, dim, L) in `dtype` (float64 = truth, float32 = timed CPU baseline)."""
    suf = "f64" if dtype == np.float64 else "f32"
    u, delta, A, B, C, D, delta_bias = (_prep(x, dtype) for x in (u, delta, A, B, C, D, delta_bias))
    b, dim, L = u.shape
    G, N = B.shape[1], B.shape[2]
    assert N <= 64 and dim % G == 0
    out = np.empty_like(u)
    fn = getattr(_load(), f"sscan_fwd_{suf}")
    fn.restype = None
    fn(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(delta_bias), _p(out),
       b, dim, L, N, G, int(bool(delta_softplus)))
    return out


def selective_scan_bwd(u, delta, A, B, C, D, delta_bias, dout, delta_softplus=True,
                       dtype=np.float64):
    """Returns (du, ddelta, dA, dB, dC, dD, ddelta_bias) -- csms6s.py:920-923 order."""
    suf = "f64" if dtype == np.float64 else "f32"
    u, delta, A, B, C, D, delta_bias, dout = (
        _prep(x, dtype) for x in (u, delta, A, B, C, D, delta_bias, dout))
    b, dim, L = u.shape
    G, N = B.shape[1], B.shape[2]
    du, dd = np.empty_like(u), np.empty_like(u)
    dA = np.empty((dim, N), dtype)
    dB, dC = np.empty_like(B), np.empty_like(C)
    dD, dbias = np.empty(dim, dtype), np.empty(dim, dtype)
    fn = getattr(_load(), f"sscan_bwd_{suf}")
    fn.restype = None
    fn(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(delta_bias), _p(dout), _p(du), _p(dd),
       _p(dA), _p(dB), _p(dC), _p(dD), _p(dbias), b, dim, L, N, G, int(bool(delta_softplus)))
    return du, dd, dA, dB, dC, (dD if D is not None else None), \
        (dbias if delta_bias is not None else None)


def selective_scan_fwd_numpy(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=True):
    """Pure numpy loop over L (small cases only): the same statement, independent code."""
    u, delta, A, B, C = (np.asarray(x, np.float64) for x in (u, delta, A, B, C))
    b, dim, L = u.shape
    G, N = B.shape[1], B.shape[2]
    x = delta + (0 if delta_bias is None else np.asarray(delta_bias, np.float64)[None, :, None])
    dl = np.where(x <= 20, np.log1p(np.exp(np.minimum(x, 20))), x) if delta_softplus else x
    Bf = np.repeat(B, dim // G, axis=1)  # (b, dim, N, L)
    Cf = np.repeat(C, dim // G, axis=1)
    h = np.zeros((b, dim, N))
    out = np.empty_like(u)
    for t in range(L):
        h = np.exp(dl[:, :, t, None] * A[None]) * h + dl[:, :, t, None] * Bf[:, :, :, t] * u[:, :, t, None]
        out[:, :, t] = (Cf[:, :, :, t] * h).sum(-1)
    if D is not None:
        out = out + u * np.asarray(D, np.float64)[None, :, None]
    return out
