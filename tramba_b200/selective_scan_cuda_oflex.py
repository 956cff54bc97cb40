"""Drop-in for the un-vendored extension module the reference imports by bare name
(`import selective_scan_cuda_oflex`, Models/SS2D/csms6s.py:741-745) and calls as

    out, x, *rest = selective_scan_cuda_oflex.fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, 1, oflex)
    du, ddelta, dA, dB, dC, dD, ddelta_bias, *rest = selective_scan_cuda_oflex.bwd(
        u, delta, A, B, C, D, delta_bias, dout, x, delta_softplus, 1)        (csms6s.py:910, 920-922)

`tramba_b200.install_extension()` registers this module in sys.modules under that name, after
which the reference's own SelectiveScanOflex runs on the sm_100a kernels unmodified.
Tensors are torch-owned; the work happens in libtramba_b200.so on the current CUDA stream.
`x` is an opaque checkpoint tensor (chunk-boundary states), as in the upstream extension.
"""
import ctypes

import torch

from . import _lib

_DTYPE = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _code(t):
    try:
        return _DTYPE[t.dtype]
    except KeyError:
        raise RuntimeError(f"selective_scan_cuda_oflex: unsupported dtype {t.dtype}")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _al(t):
    """contiguous + 16-byte aligned base pointer (a sliced view may be neither)."""
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


def _prep(u, delta, A, B, C, D, delta_bias):
    if not u.is_cuda:
        raise RuntimeError("selective_scan_cuda_oflex: CUDA tensors required (no CPU path)")
    if u.dim() != 3:
        raise RuntimeError("selective_scan_cuda_oflex: u must be (batch, dim, L)")
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    batch, dim, L = u.shape
    G, N = B.shape[1], B.shape[2]
    if delta.shape != u.shape or C.shape != B.shape or B.shape[0] != batch or B.shape[3] != L:
        raise RuntimeError("selective_scan_cuda_oflex: inconsistent shapes")
    if A.shape != (dim, N):
        raise RuntimeError("selective_scan_cuda_oflex: A must be (dim, d_state)")
    u, delta = _al(u), _al(delta.to(u.dtype))
    B, C = _al(B.to(u.dtype)), _al(C.to(u.dtype))
    A = A.contiguous().float()
    D = None if D is None else D.contiguous().float()
    delta_bias = None if delta_bias is None else delta_bias.contiguous().float()
    return u, delta, A, B, C, D, delta_bias, (batch, dim, L, N, G)


def fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows, oflex=True):
    u, delta, A, B, C, D, delta_bias, (batch, dim, L, N, G) = _prep(u, delta, A, B, C, D, delta_bias)
    in_dt = _code(u)
    out = torch.empty((batch, dim, L), device=u.device, dtype=torch.float32 if oflex else u.dtype)
    x = torch.empty(max(1, _lib.lib.tramba_sscan_ckpt_elems(batch, dim, L, N, in_dt)),
                    device=u.device, dtype=torch.float32)
    # algorithmic bytes: read u, delta, B, C; write out
    nbytes = u.numel() * (2 * u.element_size() + out.element_size()) + 2 * B.numel() * B.element_size()
    with torch.cuda.device(u.device), _lib.timed("sscan_fwd", nbytes):
        _lib.check(_lib.lib.tramba_sscan_fwd(
            u.data_ptr(), delta.data_ptr(), A.data_ptr(), B.data_ptr(), C.data_ptr(), _ptr(D), _ptr(delta_bias),
            out.data_ptr(), x.data_ptr(), batch, dim, L, N, G, in_dt, _code(out), int(bool(delta_softplus)),
            _stream()))
    return [out, x]


def bwd(u, delta, A, B, C, D, delta_bias, dout, x, delta_softplus, nrows):
    b_shape, c_shape = B.shape, C.shape
    u, delta, A, B, C, D, delta_bias, (batch, dim, L, N, G) = _prep(u, delta, A, B, C, D, delta_bias)
    in_dt = _code(u)
    if dout.dtype not in (torch.float32, u.dtype):
        dout = dout.float()
    dout = _al(dout)
    du, ddelta = torch.empty_like(u), torch.empty_like(delta)
    dA = torch.empty_like(A)
    dB = torch.empty(B.shape, device=u.device, dtype=torch.float32)
    dC = torch.empty(C.shape, device=u.device, dtype=torch.float32)
    dD = None if D is None else torch.empty_like(D)
    dbias = None if delta_bias is None else torch.empty_like(delta_bias)
    # algorithmic bytes: read u, delta, dout, B, C; write du, ddelta, dB, dC (fp32)
    nbytes = (u.numel() * (4 * u.element_size() + dout.element_size())
              + 2 * B.numel() * (B.element_size() + 4))
    with torch.cuda.device(u.device), _lib.timed("sscan_bwd", nbytes):
        _lib.check(_lib.lib.tramba_sscan_bwd(
            u.data_ptr(), delta.data_ptr(), A.data_ptr(), B.data_ptr(), C.data_ptr(), _ptr(D), _ptr(delta_bias),
            dout.data_ptr(), x.data_ptr(), du.data_ptr(), ddelta.data_ptr(), dA.data_ptr(), dB.data_ptr(),
            dC.data_ptr(), _ptr(dD), _ptr(dbias), batch, dim, L, N, G, in_dt, _code(dout),
            int(bool(delta_softplus)), _stream()))
    if B.dtype != torch.float32:
        dB, dC = dB.to(B.dtype), dC.to(C.dtype)
    return [du, ddelta, dA, dB.view(b_shape), dC.view(c_shape), dD, dbias]
