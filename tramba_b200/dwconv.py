"""Depthwise k x k convolution (+ bias, + SiLU) on the sm_100a kernel.

    y = dwconv2d(x, weight, bias=None, act="none" | "silu")

Same function as `act(conv2d(x))` with nn.Conv2d(C, C, k, padding=(k-1)//2, groups=C) -- the
depthwise conv + SiLU that produces the SS2D core's input (Models/vmamba.py:72-79, 281-285).
One kernel forward (tramba_dwconv2d_fwd), one backward (tramba_dwconv2d_bwd: dgrad + wgrad + dbias,
pre-activation recomputed).  fp32 accumulation; weight / bias are read as fp32.
"""
import ctypes

import torch

from . import _lib

_DTYPE = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}
_ACT = {"none": _lib.ACT_NONE, None: _lib.ACT_NONE, "silu": _lib.ACT_SILU}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def supported(x, weight):
    if not (x.is_cuda and x.dim() == 4 and x.dtype in _DTYPE and weight.dim() == 4):
        return False
    B, C, H, W = x.shape
    k = weight.shape[-1]
    return (weight.shape[0] == C and weight.shape[1] == 1 and weight.shape[2] == k
            and bool(_lib.lib.tramba_dwconv2d_supported(B, C, H, W, k)))


def _c16(t):
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()


class DWConv2dFn(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, bias, act):
        B, C, H, W = x.shape
        k = weight.shape[-1]
        x = _c16(x)
        w = weight.detach().contiguous().float()
        b = None if bias is None else bias.detach().contiguous().float()
        y = torch.empty_like(x)
        with torch.cuda.device(x.device), _lib.timed(f"dwconv{k}_fwd", 2 * x.numel() * x.element_size()):
            _lib.check(_lib.lib.tramba_dwconv2d_fwd(x.data_ptr(), w.data_ptr(), None if b is None else b.data_ptr(),
                                                    y.data_ptr(), B, C, H, W, k, _ACT[act], _DTYPE[x.dtype], _stream()))
        ctx.save_for_backward(x, w, b)
        ctx.meta = (act, weight.dtype, None if bias is None else bias.dtype, tuple(weight.shape))
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, w, b = ctx.saved_tensors
        act, wdt, bdt, wshape = ctx.meta
        B, C, H, W = x.shape
        k = wshape[-1]
        dy = _c16(dy.to(x.dtype))
        dx = torch.empty_like(x)
        part = torch.empty((B, C, k * k + 1), device=x.device, dtype=torch.float32)   # taps | bias, per plane
        with torch.cuda.device(x.device), _lib.timed(f"dwconv{k}_bwd", 3 * x.numel() * x.element_size()):
            _lib.check(_lib.lib.tramba_dwconv2d_bwd(
                x.data_ptr(), w.data_ptr(), None if b is None else b.data_ptr(), dy.data_ptr(), dx.data_ptr(),
                part.data_ptr(), B, C, H, W, k, _ACT[act], _DTYPE[x.dtype], _stream()))
        from .fused import _batch_sum
        part = _batch_sum(part)                                                       # fold over the batch (library kernel)
        dw = part[:, :k * k].reshape(wshape).to(wdt)
        return dx, dw, None if b is None else part[:, k * k].to(bdt), None


def dwconv2d(x, weight, bias=None, act="none"):
    """x (B, C, H, W); weight (C, 1, k, k); stride 1, padding (k-1)//2."""
    return DWConv2dFn.apply(x, weight, bias, act)
