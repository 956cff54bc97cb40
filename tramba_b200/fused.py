"""Fused SS2D core: CrossScan* -> SelectiveScanOflex -> CrossMerge* as ONE autograd operator
(Models/vmamba.py:231-259 without the permuted copies xs / ys of the feature map).

    y = ss2d_core_fused(x, dts, Bs, Cs, A, Ds, delta_bias, order, param)

    x   (B, D, H, W) image order      dts (B, K, D, L), Bs, Cs (B, K, L) scan order (as the reference's
    projections produce them)         A, Ds, delta_bias (K*D) fp32
    y   (B, D, L) fp32, image order -- what CrossMerge*.apply(SelectiveScanOflex.apply(...)) returns

Forward = one kernel (tramba_ss2d_fused_fwd); backward = one kernel (tramba_ss2d_fused_bwd) that returns dx
already merged over the K directions (CrossScan*.backward folded in: no du tensor in HBM).  N = 1 only; shapes the
fused kernels do not cover are reported by `supported()` and run un-fused.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

_DTYPE = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}
_idx_cache = {}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def order_table16_host(order, H, param=0):
    K = _lib.lib.tramba_order_num_directions(order)
    _lib.check(min(K, 0))
    t = np.empty((K, H * H), dtype=np.uint16)
    _lib.check(_lib.lib.tramba_order_table16(order, H, H, param, t.ctypes.data))
    return t


def _idx16(order, H, param, device):
    """Device table the kernels consume: the SKEWED shared-memory position of every token's pixel."""
    key = (order, H, param, device)
    hit = _idx_cache.get(key)
    if hit is None:
        K = _lib.lib.tramba_order_num_directions(order)
        t = np.empty((K, H * H), dtype=np.uint16)
        _lib.check(_lib.lib.tramba_order_table16_skewed(order, H, H, param, t.ctypes.data))
        hit = torch.from_numpy(t.view(np.int16)).to(device)
        _idx_cache[key] = hit
    return hit


def token_index(order, H, param, device):
    """(K, L) int64 pixel index of every token (for gathering small per-token tensors into scan order)."""
    key = ("i64", order, H, param, device)
    hit = _idx_cache.get(key)
    if hit is None:
        hit = torch.from_numpy(order_table16_host(order, H, param).astype(np.int64)).to(device)
        _idx_cache[key] = hit
    return hit


def _sched(order, H, param, device):
    """(device (K, len) int32 schedule, host (K,) int32 kinds) of tramba_order_sched for this order / size."""
    key = ("sched", order, H, param, device)
    hit = _idx_cache.get(key)
    if hit is None:
        K = _lib.lib.tramba_order_num_directions(order)
        n = int(_lib.lib.tramba_order_sched_len(order, H, H, param))
        if n == 0:
            raise RuntimeError("tramba_b200: no permutation schedule for this order / size")
        t = np.empty((K, n), dtype=np.uint32)
        kinds = np.empty((2 * K,), dtype=np.int32)       # K kinds, then K used lengths
        _lib.check(_lib.lib.tramba_order_sched(order, H, H, param, t.ctypes.data, kinds.ctypes.data))
        hit = (torch.from_numpy(t.view(np.int32)).to(device), kinds)
        _idx_cache[key] = hit
    return hit


def supported(order, K, D, H, W, N, dtype, need_bwd=True):
    code = _DTYPE.get(dtype)
    if code is None or H != W:
        return False
    return bool(_lib.lib.tramba_ss2d_fused_supported(order, K, D, H, W, N, code, int(need_bwd)))


def _ptr(t):
    return None if t is None else t.data_ptr()


def _c16(t):
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()


class SS2DCoreFused(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, dts, Bs, Cs, A, Ds, delta_bias, order, param, delta_softplus=True):
        B, D, H, W = x.shape
        K = dts.shape[1]
        L = H * W
        dt = x.dtype
        x, dts = _c16(x), _c16(dts.to(dt))
        Bs, Cs = _c16(Bs.to(dt)), _c16(Cs.to(dt))
        A = A.contiguous().float()
        Ds = None if Ds is None else Ds.contiguous().float()
        delta_bias = None if delta_bias is None else delta_bias.contiguous().float()
        idx = _idx16(order, H, param, x.device)
        y = torch.empty((B, D, L), device=x.device, dtype=torch.float32)
        es = x.element_size()
        nbytes = (B * D * L) * (es + 4) + dts.numel() * es + 2 * Bs.numel() * es   # fused core
        with torch.cuda.device(x.device), _lib.timed("ss2d_fused_fwd", nbytes):
            _lib.check(_lib.lib.tramba_ss2d_fused_fwd(
                x.data_ptr(), dts.data_ptr(), Bs.data_ptr(), Cs.data_ptr(), A.data_ptr(), _ptr(Ds),
                _ptr(delta_bias), idx.data_ptr(), y.data_ptr(), order, K, B, D, H, W, 1, _DTYPE[dt],
                int(bool(delta_softplus)), _stream()))
        ctx.save_for_backward(x, dts, Bs, Cs, A, Ds, delta_bias, idx)
        ctx.meta = (order, param, K, B, D, H, W, bool(delta_softplus))
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        """One kernel: CrossMerge*.backward -> SelectiveScanOflex.backward -> CrossScan*.backward (csms6s.py:45-55,
        914-923, 25-31).  dx comes back merged over the K directions in fp32; du never exists in HBM."""
        x, dts, Bs, Cs, A, Ds, delta_bias, idx = ctx.saved_tensors
        order, param, K, B, D, H, W, softplus = ctx.meta
        L = H * W
        dt = x.dtype
        dy = _c16(dy.float())
        sched, kinds = _sched(order, H, param, x.device)
        # all fp32 accumulation targets from ONE workspace (-> one memset inside the entry point)
        n_dx, n_bc, n_row = B * D * L, B * K * L, K * D
        ws = torch.empty((n_dx + 2 * n_bc + 3 * n_row,), device=x.device, dtype=torch.float32)
        dx, dBs, dCs, dA, dD, db = torch.split(ws, [n_dx, n_bc, n_bc, n_row, n_row, n_row])
        ddts = torch.empty((B, K, D, L), device=x.device, dtype=dt)
        es = x.element_size()
        # SURVEY 8(d) operator formula per direction (read u, delta, dout; write du, ddelta; B/C in, dB/dC out): what the
        # kernel REPLACES.  bytes_min = what it must move itself (x, dy once; dx once; delta / ddelta per direction).
        nbytes = K * B * D * L * (4 * es + 4) + 2 * Bs.numel() * (es + 4)
        with torch.cuda.device(x.device), _lib.timed("ss2d_fused_bwd", nbytes):
            _lib.check(_lib.lib.tramba_ss2d_fused_bwd(
                x.data_ptr(), dts.data_ptr(), Bs.data_ptr(), Cs.data_ptr(), A.data_ptr(), _ptr(Ds), _ptr(delta_bias),
                sched.data_ptr(), sched.shape[1], kinds.ctypes.data, dy.data_ptr(), dx.data_ptr(), ddts.data_ptr(),
                dBs.data_ptr(), dCs.data_ptr(), dA.data_ptr(), None if Ds is None else dD.data_ptr(),
                None if delta_bias is None else db.data_ptr(), order, K, B, D, H, W, 1, _DTYPE[dt],
                int(softplus), _stream()))
        return (dx.to(dt).view(B, D, H, W), ddts, dBs.view(B, K, L).to(dt), dCs.view(B, K, L).to(dt), dA,
                None if Ds is None else dD, None if delta_bias is None else db, None, None, None)


def ss2d_core_fused(x, dts, Bs, Cs, A, Ds, delta_bias, order, param=0, delta_softplus=True):
    return SS2DCoreFused.apply(x, dts, Bs, Cs, A, Ds, delta_bias, order, param, delta_softplus)


class ProjGather(torch.autograd.Function):
    """src (B, K, R+2N, L) image order -> (dts (B, K, R, L), Bs (B, K, N, L), Cs (B, K, N, L)) scan order: the
    projected channels of the reference's xs -> x_dbl -> split (Models/vmamba.py:231-235) without xs.
    One kernel each way (tramba_proj_gather / tramba_proj_scatter)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, src, order, param, H, R, N):
        from .ops import _order_plan
        B, K, CX, L = src.shape
        src = _c16(src)
        _, plan = _order_plan(order, H, param, src.device)
        dts = torch.empty((B, K, R, L), device=src.device, dtype=src.dtype)
        Bs = torch.empty((B, K, N, L), device=src.device, dtype=src.dtype)
        Cs = torch.empty((B, K, N, L), device=src.device, dtype=src.dtype)
        with torch.cuda.device(src.device), _lib.timed("proj_gather", 2 * src.numel() * src.element_size()):
            _lib.check(_lib.lib.tramba_proj_gather(src.data_ptr(), dts.data_ptr(), Bs.data_ptr(), Cs.data_ptr(),
                                                   plan.data_ptr(), K, B, R, N, L, _DTYPE[src.dtype], _stream()))
        ctx.plan = plan
        ctx.meta = (B, K, R, N, L, src.dtype)
        return dts, Bs, Cs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gdts, gBs, gCs):
        B, K, R, N, L, dt = ctx.meta
        gdts = _c16(gdts.to(dt))
        bc = torch.float32 if (gBs.dtype == torch.float32 and gCs.dtype == torch.float32) else dt
        gBs, gCs = _c16(gBs.to(bc)), _c16(gCs.to(bc))
        g = torch.empty((B, K, R + 2 * N, L), device=gdts.device, dtype=dt)
        with torch.cuda.device(gdts.device), _lib.timed("proj_scatter", 2 * g.numel() * g.element_size()):
            _lib.check(_lib.lib.tramba_proj_scatter(gdts.data_ptr(), gBs.data_ptr(), gCs.data_ptr(), g.data_ptr(),
                                                    ctx.plan.data_ptr(), K, B, R, N, L, _DTYPE[dt], _DTYPE[bc],
                                                    _stream()))
        return g, None, None, None, None, None


def proj_gather(src, order, param, H, R, N):
    return ProjGather.apply(src, order, param, H, R, N)


def _batch_sum(t, groups=0):
    """(B, ...) -> (...): sum over the leading (image batch) axis on the library's kernel (fp32 accumulation).
    groups = G > 0: (G, B, ...) -> (G, ...) in one launch."""
    t = _c16(t)
    G = max(groups, 1)
    assert groups == 0 or t.shape[0] == groups
    out = torch.empty(t.shape[2:] if groups else t.shape[1:], device=t.device, dtype=t.dtype)
    n = out.numel()
    if groups:
        out = torch.empty((G,) + tuple(out.shape), device=t.device, dtype=t.dtype)
    with torch.cuda.device(t.device), _lib.timed("batch_sum", t.numel() * t.element_size()):
        _lib.check(_lib.lib.tramba_batch_sum(t.data_ptr(), out.data_ptr(), G, t.shape[1] if groups else t.shape[0], n,
                                             _DTYPE[t.dtype], _stream()))
    return out


class ProjX(torch.autograd.Function):
    """x_proj on the image-ordered map as ONE GEMM per image: Y (B, M, L) = W (M, D) @ X (B, D, L)
    (Models/vmamba.py:233-234 evaluated before the gather).  Same products as torch.bmm on the expanded
    weight; the weight gradient sum_b dY_b X_b^T is one batched GEMM + the library's batch reduction instead of
    autograd's generic reduce over the expanded dimension."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=None)
    def forward(ctx, W, X):
        if torch.is_autocast_enabled():
            dt = torch.get_autocast_dtype("cuda")
            W, X = W.to(dt), X.to(dt)
        elif W.dtype != X.dtype:            # fp32 master weight on lower-precision activations: follow the activations
            W = W.to(X.dtype)
        ctx.save_for_backward(W, X)
        with torch.autocast("cuda", enabled=False):
            return torch.bmm(W.unsqueeze(0).expand(X.shape[0], -1, -1), X)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dY):
        W, X = ctx.saved_tensors
        dY = dY.to(X.dtype)
        B = X.shape[0]
        dX = torch.bmm(W.t().unsqueeze(0).expand(B, -1, -1), dY) if ctx.needs_input_grad[1] else None
        dW = _batch_sum(torch.bmm(dY, X.transpose(1, 2))) if ctx.needs_input_grad[0] else None
        return dW, dX


class ProjDt(torch.autograd.Function):
    """dt_proj per direction: Y (B, K, D, L) = W (K, D, R) @ X (B, K, R, L) (Models/vmamba.py:236); weight gradient as
    in ProjX."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=None)
    def forward(ctx, W, X):
        if torch.is_autocast_enabled():
            dt = torch.get_autocast_dtype("cuda")
            W, X = W.to(dt), X.to(dt)
        elif W.dtype != X.dtype:            # fp32 master weight on lower-precision activations: follow the activations
            W = W.to(X.dtype)
        ctx.save_for_backward(W, X)
        with torch.autocast("cuda", enabled=False):
            return torch.matmul(W, X)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dY):
        W, X = ctx.saved_tensors
        dY = dY.to(X.dtype)
        dX = torch.matmul(W.transpose(1, 2), dY) if ctx.needs_input_grad[1] else None
        dW = _batch_sum(torch.matmul(dY, X.transpose(2, 3))) if ctx.needs_input_grad[0] else None
        return dW, dX
