"""Operators of the 2D selective-scan hot path, with the reference's names and signatures.

    CrossScan / CrossMerge                       Models/SS2D/csms6s.py:13-55
    CrossScan_Dilation / CrossMerge_Dilation     Models/SS2D/csms6s.py:64-103
    CrossScan_Window / CrossMerge_Window         Models/SS2D/csms6s.py:113-152
    CrossScan_Line / CrossMerge_Line             Models/SS2D/csms6s.py:161-216   (Helix-SS2D, K = 8)
    SelectiveScanOflex                           Models/SS2D/csms6s.py:904-923
Ablation orders (not on the Tramba-V path; table-driven kernels, never the fused core):
    CrossScan_Line_4direction / _Spiral / _Hilbert / _Diagonal / _DS / _Ab_2direction / _Ab_1direction
    and their CrossMerge_* twins                 Models/SS2D/csms6s.py:219-737
    make_custom_pair(name, table)                any caller-supplied (K, L) gather table

Each is a torch.autograd.Function used through `.apply`, exactly like the reference classes,
so they can be passed as `scan=` / `merge=` to SS2D or assigned over the reference's own
classes.  All compute goes through the C-ABI library (tramba_b200/_lib.py);
torch only owns the memory and the stream.  No CPU path exists.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import selective_scan_cuda_oflex as _ext

# --- module-level scan hyper-parameters, as in the reference (csms6s.py:58-59, 107-108, 157)
sizes = [12, 24, 48, 96]
dilation_rate = [4, 4, 4, 4]
window_sizes = [4, 8, 12, 16]
dilation_rate_for = {str(s): r for s, r in zip(sizes, dilation_rate)}
window_size_for = {str(s): w for s, w in zip(sizes, window_sizes)}
line_sizes = [12, 24, 48, 96, 7, 14, 28, 56]
# build-side extension for the 768x768 stress configuration: H = 192
dilation_rate_for["192"] = 4
window_size_for["192"] = 32
line_sizes.append(192)

_DTYPE = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}
_plan_cache = {}


def _dtype_code(t):
    try:
        return _DTYPE[t.dtype]
    except KeyError:
        raise TypeError(f"tramba_b200: unsupported dtype {t.dtype} (float32 and bfloat16 only)")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t):
    if not t.is_cuda:
        raise RuntimeError("tramba_b200 has no CPU path: tensors must live on an sm_100 CUDA device")


def order_plan_host(order, H, param=0):
    """The order plan (forward table + inverse CSR) as a numpy int32 array; host only."""
    K = _lib.lib.tramba_order_num_directions(order)
    _lib.check(min(K, 0))
    L = H * H
    plan = np.empty(_lib.lib.tramba_order_plan_elems(K, L), dtype=np.int32)
    _lib.check(_lib.lib.tramba_order_plan(order, H, H, param, plan.ctypes.data))
    return K, plan


def order_table_host(order, H, param=0):
    K, plan = order_plan_host(order, H, param)
    return plan[: K * H * H].reshape(K, H * H)


def _order_plan(order, H, param, device):
    key = (order, H, param, device)
    hit = _plan_cache.get(key)
    if hit is None:
        K, plan = order_plan_host(order, H, param)
        hit = (K, torch.from_numpy(plan).to(device))
        _plan_cache[key] = hit
    return hit


def _check_square(H, W):
    if H != W:
        raise ValueError("tramba_b200: scan orders are defined for square maps only (as in the reference)")


def cross_scan(x, order, param=0):
    """x: (B, C, H, W) -> xs: (B, K, C, H*W)."""
    _require_cuda(x)
    B, C, H, W = x.shape
    _check_square(H, W)
    x = x.contiguous()
    K, plan = _order_plan(order, H, param, x.device)
    xs = x.new_empty((B, K, C, H * W))
    with torch.cuda.device(x.device), _lib.timed("cross_scan", (1 + K) * x.numel() * x.element_size()):
        _lib.check(_lib.lib.tramba_cross_scan(x.data_ptr(), xs.data_ptr(), plan.data_ptr(), order, K, B, C, H, W,
                                              _dtype_code(x), _stream()))
    return xs


def cross_merge(ys, order, H, W, param=0):
    """ys: (B, K, D, L) (any view of it) -> y: (B, D, L)."""
    _require_cuda(ys)
    _check_square(H, W)
    B, K, D = ys.shape[:3]
    ys = ys.contiguous()
    K2, plan = _order_plan(order, H, param, ys.device)
    if K2 != K:
        raise ValueError(f"tramba_b200: order expects K={K2} directions, got {K}")
    y = ys.new_empty((B, D, H * W))
    with torch.cuda.device(ys.device), _lib.timed("cross_merge", (1 + K) * y.numel() * y.element_size()):
        _lib.check(_lib.lib.tramba_cross_merge(ys.data_ptr(), y.data_ptr(), plan.data_ptr(), order, K, B, D, H, W,
                                               _dtype_code(ys), _stream()))
    return y


def _custom_plan(table, device):
    """Order plan (device) of a caller-supplied (K, L) gather table; cached on the table's bytes."""
    table = np.ascontiguousarray(table, dtype=np.int32)
    if table.ndim != 2:
        raise ValueError("tramba_b200: a custom order table must be (K, L)")
    key = ("custom", table.shape, hash(table.tobytes()), device)
    hit = _plan_cache.get(key)
    if hit is None:
        K, L = table.shape
        plan = np.empty(_lib.lib.tramba_order_plan_elems(K, L), dtype=np.int32)
        _lib.check(_lib.lib.tramba_order_plan_from_table(K, L, table.ctypes.data, plan.ctypes.data))
        hit = (K, torch.from_numpy(plan).to(device))
        _plan_cache[key] = hit
    return hit


def _make_pair(name, order, param_of):
    """Build the CrossScan*/CrossMerge* pair for one order family."""

    class Scan(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            B, C, H, W = x.shape
            ctx.shape = (B, C, H, W)
            return cross_scan(x, order, param_of(H))

        @staticmethod
        def backward(ctx, ys):
            B, C, H, W = ctx.shape
            return cross_merge(ys, order, H, W, param_of(H)).view(B, -1, H, W)

    class Merge(torch.autograd.Function):
        @staticmethod
        def forward(ctx, ys):
            B, K, D, H, W = ys.shape
            ctx.shape = (H, W)
            return cross_merge(ys.view(B, K, D, -1), order, H, W, param_of(H))

        @staticmethod
        def backward(ctx, x):
            H, W = ctx.shape
            B, C, L = x.shape
            xs = cross_scan(x.view(B, C, H, W), order, param_of(H))
            return xs.view(B, -1, C, H, W)

    Scan.__name__ = Scan.__qualname__ = "CrossScan" + name
    Merge.__name__ = Merge.__qualname__ = "CrossMerge" + name
    Scan.tramba_order = Merge.tramba_order = order            # lets SS2D route the pair to the fused core
    Scan.tramba_param = Merge.tramba_param = staticmethod(param_of)
    return Scan, Merge


def _line_param(H):
    if H not in line_sizes:
        raise KeyError(str(H))  # reference: spiral_line_index[str(H)] (csms6s.py:166)
    return 0


CrossScan, CrossMerge = _make_pair("", _lib.ORDER_STD, lambda H: 0)
CrossScan_Dilation, CrossMerge_Dilation = _make_pair("_Dilation", _lib.ORDER_DILATION,
                                                     lambda H: dilation_rate_for[str(H)])
CrossScan_Window, CrossMerge_Window = _make_pair("_Window", _lib.ORDER_WINDOW, lambda H: window_size_for[str(H)])
CrossScan_Line, CrossMerge_Line = _make_pair("_Line", _lib.ORDER_LINE, _line_param)


# --- the reference's ablation orders (csms6s.py:219-737).  They are not on the Tramba-V path: the pairs carry
# no `tramba_order` routing attribute, so SS2D always runs them op-by-op on the table-driven kernels.
def _ref_sizes_only(H):
    assert H in sizes            # reference: `assert H == W and H in [12, 24, 48, 96]` (csms6s.py:272,392)
    return 0


def _ablation_pair(name, order, param_of):
    S, M = _make_pair(name, order, param_of)
    for cls in (S, M):
        cls.tramba_ablation_order = order
        del cls.tramba_order, cls.tramba_param
    return S, M


CrossScan_Line_4direction, CrossMerge_Line_4direction = _ablation_pair("_Line_4direction", _lib.ORDER_LINE4, _line_param)
CrossScan_Spiral, CrossMerge_Spiral = _ablation_pair("_Spiral", _lib.ORDER_SPIRAL, _ref_sizes_only)
CrossScan_Hilbert, CrossMerge_Hilbert = _ablation_pair("_Hilbert", _lib.ORDER_HILBERT, _ref_sizes_only)
CrossScan_Diagonal, CrossMerge_Diagonal = _ablation_pair("_Diagonal", _lib.ORDER_DIAGONAL, lambda H: 0)
CrossScan_DS, CrossMerge_DS = _ablation_pair("_DS", _lib.ORDER_DS, _line_param)
CrossScan_Ab_2direction, CrossMerge_Ab_2direction = _ablation_pair("_Ab_2direction", _lib.ORDER_AB2, lambda H: 0)
CrossScan_Ab_1direction, CrossMerge_Ab_1direction = _ablation_pair("_Ab_1direction", _lib.ORDER_AB1, lambda H: 0)


def make_custom_pair(name, table_of):
    """(CrossScan, CrossMerge) autograd pair for ANY gather table: `table_of(H)` returns the (K, L = H*H) flat
    pixel index of every token (repeats allowed, like the helix lists).  scan: xs[b,k,c,t] = x[b,c].flat[T[k,t]];
    merge: its exact transpose, summed sequentially over k then t (deterministic, no atomics)."""

    plans = {}                       # (H, device) -> (K, device plan): table_of() runs once per map size

    def plan_for(H, device):
        hit = plans.get((H, device))
        if hit is None:
            hit = plans[(H, device)] = _custom_plan(table_of(H), device)
        return hit

    def scan(x):
        B, C, H, W = x.shape
        _check_square(H, W)
        _require_cuda(x)
        x = x.contiguous()
        K, plan = plan_for(H, x.device)
        xs = x.new_empty((B, K, C, H * W))
        with torch.cuda.device(x.device), _lib.timed("cross_scan", (1 + K) * x.numel() * x.element_size()):
            _lib.check(_lib.lib.tramba_cross_scan(x.data_ptr(), xs.data_ptr(), plan.data_ptr(), _lib.ORDER_CUSTOM, K,
                                                  B, C, H, W, _dtype_code(x), _stream()))
        return xs

    def merge(ys, H, W):
        _check_square(H, W)
        _require_cuda(ys)
        B, K, D = ys.shape[:3]
        ys = ys.contiguous()
        K2, plan = plan_for(H, ys.device)
        if K2 != K:
            raise ValueError(f"tramba_b200: order expects K={K2} directions, got {K}")
        y = ys.new_empty((B, D, H * W))
        with torch.cuda.device(ys.device), _lib.timed("cross_merge", (1 + K) * y.numel() * y.element_size()):
            _lib.check(_lib.lib.tramba_cross_merge(ys.data_ptr(), y.data_ptr(), plan.data_ptr(), _lib.ORDER_CUSTOM, K,
                                                   B, D, H, W, _dtype_code(ys), _stream()))
        return y

    class Scan(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            ctx.shape = x.shape
            return scan(x)

        @staticmethod
        def backward(ctx, ys):
            B, C, H, W = ctx.shape
            return merge(ys, H, W).view(B, -1, H, W)

    class Merge(torch.autograd.Function):
        @staticmethod
        def forward(ctx, ys):
            B, K, D, H, W = ys.shape
            ctx.shape = (H, W)
            return merge(ys.view(B, K, D, -1), H, W)

        @staticmethod
        def backward(ctx, x):
            H, W = ctx.shape
            B, C, L = x.shape
            return scan(x.view(B, C, H, W)).view(B, -1, C, H, W)

    Scan.__name__ = Scan.__qualname__ = "CrossScan" + name
    Merge.__name__ = Merge.__qualname__ = "CrossMerge" + name
    return Scan, Merge


class SelectiveScanOflex(torch.autograd.Function):
    """Same contract as the reference class (csms6s.py:904-923): autocast is disabled inside,
    inputs arrive in whatever dtype autocast produced, `nrows`/`backnrows` are accepted and
    ignored (the reference sends 1 to the extension), backward returns 11 values."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1,
                oflex=True):
        ctx.delta_softplus = delta_softplus
        out, x, *rest = _ext.fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, 1, oflex)
        ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout, *args):
        u, delta, A, B, C, D, delta_bias, x = ctx.saved_tensors
        if dout.stride(-1) != 1:
            dout = dout.contiguous()
        du, ddelta, dA, dB, dC, dD, ddelta_bias, *rest = _ext.bwd(
            u, delta, A, B, C, D, delta_bias, dout, x, ctx.delta_softplus, 1)
        return (du, ddelta, dA, dB, dC, dD, ddelta_bias, None, None, None, None)


def selective_scan_fn(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=True, oflex=True):
    """The closure the reference's SS2D builds at Models/vmamba.py:147-148."""
    return SelectiveScanOflex.apply(u, delta, A, B, C, D, delta_bias, delta_softplus, -1, -1, oflex)
