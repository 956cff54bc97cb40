"""LayerNorm2d (+ GELU) on the sm_100a kernel: layer norm over the channel axis of an NCHW map.

    y = layer_norm2d(x, weight, bias, eps=1e-5, act="none" | "gelu", out_dtype=None)

Same function as LayerNorm2d.forward (Models/modules.py:22-27) -- permute(0,2,3,1) -> F.layer_norm
-> permute(0,3,1,2) -- and, with act="gelu", as the `out_act(out_norm(y).to(dtype))` chain that
follows the cross-merge in SS2D (Models/vmamba.py:271,287), evaluated in ONE pass over the map
(tramba_layernorm2d_fwd / _bwd); fp32 statistics and arithmetic, output rounded once to out_dtype.
"""
import ctypes

import torch

from . import _lib

_DTYPE = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}
_ACT = {"none": _lib.ACT_NONE, None: _lib.ACT_NONE, "gelu": _lib.ACT_GELU}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def supported(x, out_dtype=None):
    return (x.is_cuda and x.dim() >= 3 and x.dtype in _DTYPE and (out_dtype or x.dtype) in _DTYPE
            and x.shape[0] <= 65535 and x.numel() > 0)


def _c8(t):
    t = t.contiguous()
    return t if t.data_ptr() % 8 == 0 else t.clone()


class LayerNorm2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, act, out_dtype):
        shape = x.shape
        B, C = shape[0], shape[1]
        L = x.numel() // (B * C)
        x = _c8(x)
        w = weight.detach().contiguous().float()
        b = None if bias is None else bias.detach().contiguous().float()
        y = torch.empty(shape, device=x.device, dtype=out_dtype)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or (bias is not None and ctx.needs_input_grad[2])
        mean = torch.empty((B, L), device=x.device, dtype=torch.float32) if need else None
        rstd = torch.empty((B, L), device=x.device, dtype=torch.float32) if need else None
        nbytes = x.numel() * (x.element_size() + y.element_size())
        with torch.cuda.device(x.device), _lib.timed("layernorm2d_fwd", nbytes):
            _lib.check(_lib.lib.tramba_layernorm2d_fwd(
                x.data_ptr(), w.data_ptr(), None if b is None else b.data_ptr(), y.data_ptr(),
                None if mean is None else mean.data_ptr(), None if rstd is None else rstd.data_ptr(),
                B, C, L, float(eps), _ACT[act], _DTYPE[x.dtype], _DTYPE[out_dtype], _stream()))
        ctx.save_for_backward(x, w, b, mean, rstd)
        ctx.meta = (B, C, L, act, weight.dtype, None if bias is None else bias.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, b, mean, rstd = ctx.saved_tensors
        B, C, L, act, wdt, bdt = ctx.meta
        if gy.dtype not in _DTYPE:
            gy = gy.float()
        gy = _c8(gy)
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            tiles = _lib.lib.tramba_layernorm2d_tiles(B, C, L)
            _lib.check(min(tiles, 0))
            part = torch.empty((2 if b is not None else 1, B * tiles, C), device=x.device, dtype=torch.float32)
            gw, gb = part[0], (part[1] if b is not None else None)
            nbytes = x.numel() * (2 * x.element_size() + gy.element_size())
            with _lib.timed("layernorm2d_bwd", nbytes):
                _lib.check(_lib.lib.tramba_layernorm2d_bwd(
                    gy.data_ptr(), x.data_ptr(), w.data_ptr(), None if b is None else b.data_ptr(), mean.data_ptr(),
                    rstd.data_ptr(), gx.data_ptr(), gw.data_ptr(), None if gb is None else gb.data_ptr(),
                    B, C, L, _ACT[act], _DTYPE[x.dtype], _DTYPE[gy.dtype], _stream()))
        from .fused import _batch_sum                            # fold the per-tile partials (library kernel)
        if gb is not None and (C * 4) % 16 == 0:
            gw, gb = _batch_sum(part, groups=2)                    # dweight | dbias in one launch
        else:
            gw = _batch_sum(gw)
            gb = None if gb is None else _batch_sum(gb)
        return gx, gw.to(wdt), None if gb is None else gb.to(bdt), None, None, None


def layer_norm2d(x, weight, bias=None, eps=1e-5, act="none", out_dtype=None):
    """x (B, C, *spatial) -> same shape; statistics over C.  out_dtype None = x.dtype."""
    return LayerNorm2dFn.apply(x, weight, bias, eps, act, out_dtype or x.dtype)
