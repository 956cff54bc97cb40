"""CPU emulation of libtramba_b200's COMPUTE entry points -- TEST INFRASTRUCTURE ONLY.

Purpose: exercise the product's HOST logic (tramba_b200/ops.py, fused.py, ss2d.py, norm.py, dwconv.py, dct.py,
selective_scan_cuda_oflex.py: argument marshalling, shapes, dtype handling, autograd wiring, routing between the
fused core and the op-by-op path) in the `-m "not gpu"` suite, where no kernel can run.  `install()` replaces
`tramba_b200._lib.lib` by a proxy whose compute entry points read and write the SAME raw pointers the kernels
would, but compute with the oracle (oracle/*.py, float64) on host memory; host-only entry points (order tables,
plans, *_supported) still go to the real library.  Tensors stay on the CPU: `.cuda()` is the identity,
`is_cuda` reads True and the stream / device helpers are stubbed -- all inside `installed()`'s scope only.

This is NOT a CPU fallback of the product (nothing under tramba_b200/ imports it and the product still fails
loudly without a GPU); it says nothing about the kernels -- those are checked by the `-m gpu` tests.
"""
import contextlib
import ctypes

import numpy as np
import torch

from oracle import dct as odct
from oracle import dwconv as odw
from oracle import layernorm2d as oln
from oracle import scan_orders as so
from oracle import sscan as osscan

F32, BF16 = 0, 1


class _Buf:
    """A typed window on raw host memory: .get() -> float64 ndarray, .set(values) rounds to the buffer's dtype."""

    def __init__(self, ptr, shape, code):
        self.shape, self.code = tuple(int(s) for s in shape), code
        n = int(np.prod(self.shape))
        ctype = ctypes.c_float if code == F32 else ctypes.c_uint16
        self.raw = np.ctypeslib.as_array((ctype * n).from_address(ptr)).reshape(self.shape)

    def get(self):
        if self.code == F32:
            return self.raw.astype(np.float64)
        return (self.raw.astype(np.uint32) << 16).view(np.float32).astype(np.float64)

    def set(self, val):
        val = np.ascontiguousarray(np.broadcast_to(val, self.shape), dtype=np.float32)
        if self.code == F32:
            self.raw[...] = val
        else:
            self.raw[...] = torch.from_numpy(val).bfloat16().view(torch.int16).numpy().view(np.uint16)


def _buf(ptr, shape, code=F32):
    return None if not ptr else _Buf(ptr, shape, code)


def _i32(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_int32 * n).from_address(ptr)).astype(np.int64)


def _u16(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_uint16 * n).from_address(ptr)).astype(np.int64)


def _unskew(t):                     # inverse of ss2d_fused.cuh: fused_skew(p) = p + 2*(p >> 6)
    return t - 2 * (t // 66)


_FAMILY = {0: "std", 1: "dilation", 2: "window", 3: "line"}


def _merge(ys, tab, order):
    return so.cross_merge(ys, tab, _FAMILY.get(order, "custom"))


def _get(b):
    return None if b is None else b.get()


# ------------------------------------------------------------------------------------ entry points
def tramba_cross_scan(x, xs, plan, order, K, B, C, H, W, dtype, stream):
    L = H * W
    tab = _i32(plan, K * L).reshape(K, L)
    _buf(xs, (B, K, C, L), dtype).set(so.cross_scan(_buf(x, (B, C, H, W), dtype).get(), tab))
    return 0


def tramba_cross_merge(ys, y, plan, order, K, B, D, H, W, dtype, stream):
    L = H * W
    tab = _i32(plan, K * L).reshape(K, L)
    out = _merge(_buf(ys, (B, K, D, L), dtype).get().astype(np.float32), tab, order)   # fp32 sums, like the kernel
    _buf(y, (B, D, L), dtype).set(out)
    return 0


def tramba_sscan_fwd(u, delta, A, Bm, Cm, D, bias, out, ckpt, batch, dim, L, N, groups, in_dtype, out_dtype,
                     softplus, stream):
    o = osscan.selective_scan_fwd(_buf(u, (batch, dim, L), in_dtype).get(), _buf(delta, (batch, dim, L), in_dtype).get(),
                                  _buf(A, (dim, N)).get(), _buf(Bm, (batch, groups, N, L), in_dtype).get(),
                                  _buf(Cm, (batch, groups, N, L), in_dtype).get(), _get(_buf(D, (dim,))),
                                  _get(_buf(bias, (dim,))), bool(softplus))
    _buf(out, (batch, dim, L), out_dtype).set(o)      # ckpt stays untouched: the emulated bwd recomputes
    return 0


def tramba_sscan_bwd(u, delta, A, Bm, Cm, D, bias, dout, ckpt, du, ddelta, dA, dB, dC, dD, dbias, batch, dim, L, N,
                     groups, in_dtype, dout_dtype, softplus, stream):
    g = osscan.selective_scan_bwd(_buf(u, (batch, dim, L), in_dtype).get(), _buf(delta, (batch, dim, L), in_dtype).get(),
                                  _buf(A, (dim, N)).get(), _buf(Bm, (batch, groups, N, L), in_dtype).get(),
                                  _buf(Cm, (batch, groups, N, L), in_dtype).get(), _get(_buf(D, (dim,))),
                                  _get(_buf(bias, (dim,))), _buf(dout, (batch, dim, L), dout_dtype).get(), bool(softplus))
    _buf(du, (batch, dim, L), in_dtype).set(g[0])
    _buf(ddelta, (batch, dim, L), in_dtype).set(g[1])
    _buf(dA, (dim, N)).set(g[2])
    _buf(dB, (batch, groups, N, L)).set(g[3])
    _buf(dC, (batch, groups, N, L)).set(g[4])
    if dD:
        _buf(dD, (dim,)).set(g[5])
    if dbias:
        _buf(dbias, (dim,)).set(g[6])
    return 0


def _fused_operands(x, dts, Bs, Cs, A, Dp, bias, idx16, K, B, D, H, W, dt):
    L = H * W
    tab = _unskew(_u16(idx16, K * L)).reshape(K, L)
    xs = so.cross_scan(_buf(x, (B, D, H, W), dt).get(), tab).reshape(B, K * D, L)
    return (tab, xs, _buf(dts, (B, K * D, L), dt).get(), _buf(A, (K * D, 1)).get(),
            _buf(Bs, (B, K, 1, L), dt).get(), _buf(Cs, (B, K, 1, L), dt).get(), _get(_buf(Dp, (K * D,))),
            _get(_buf(bias, (K * D,))))


def tramba_ss2d_fused_fwd(x, dts, Bs, Cs, A, Dp, bias, idx16, y, order, K, B, D, H, W, N, in_dtype, softplus, stream):
    assert N == 1
    tab, xs, d, A_, B_, C_, D_, b_ = _fused_operands(x, dts, Bs, Cs, A, Dp, bias, idx16, K, B, D, H, W, in_dtype)
    ys = osscan.selective_scan_fwd(xs, d, A_, B_, C_, D_, b_, bool(softplus)).reshape(B, K, D, H * W)
    _buf(y, (B, D, H * W)).set(so.cross_merge(ys, tab, "custom"))
    return 0


def _sched_table(sched, sched_len, K, L):
    """(K, L) pixel of every token, recovered from the permutation schedule (entry = 2*pixel | 2*token << 16)."""
    ent = np.ctypeslib.as_array((ctypes.c_uint32 * (K * sched_len)).from_address(sched)).reshape(K, sched_len)
    tab = np.full((K, L), -1, dtype=np.int64)
    for k in range(K):
        e = ent[k][ent[k] != 0xFFFFFFFF].astype(np.int64)
        tab[k, (e >> 16) >> 1] = (e & 0xFFFF) >> 1
    assert (tab >= 0).all(), "schedule does not cover every token"
    return tab


def _sched_operands(x, dts, Bs, Cs, A, Dp, bias, sched, sched_len, K, B, D, H, W, dt):
    L = H * W
    tab = _sched_table(sched, sched_len, K, L)
    xs = so.cross_scan(_buf(x, (B, D, H, W), dt).get(), tab).reshape(B, K * D, L)
    return (tab, xs, _buf(dts, (B, K * D, L), dt).get(), _buf(A, (K * D, 1)).get(),
            _buf(Bs, (B, K, 1, L), dt).get(), _buf(Cs, (B, K, 1, L), dt).get(), _get(_buf(Dp, (K * D,))),
            _get(_buf(bias, (K * D,))))


def tramba_ss2d_fused_bwd(x, dts, Bs, Cs, A, Dp, bias, sched, sched_len, kinds, dy, dx, ddts, dBs, dCs, dA, dD, dbias,
                             order, K, B, D, H, W, N, in_dtype, softplus, stream):
    assert N == 1
    L = H * W
    tab, xs, d, A_, B_, C_, D_, b_ = _sched_operands(x, dts, Bs, Cs, A, Dp, bias, sched, sched_len, K, B, D, H, W, in_dtype)
    dys = so.cross_scan(_buf(dy, (B, D, H, W)).get(), tab).reshape(B, K * D, L)     # CrossMerge.backward
    g = osscan.selective_scan_bwd(xs, d, A_, B_, C_, D_, b_, dys, bool(softplus))
    _buf(dx, (B, D, L)).set(so.cross_merge(g[0].reshape(B, K, D, L), tab, "custom"))   # CrossScan.backward, folded in
    _buf(ddts, (B, K * D, L), in_dtype).set(g[1])
    _buf(dBs, (B, K, 1, L)).set(g[3])
    _buf(dCs, (B, K, 1, L)).set(g[4])
    for ptr, val in ((dA, g[2].reshape(-1)), (dD, g[5]), (dbias, g[6])):               # totals over the batch
        if ptr:
            _buf(ptr, (K * D,)).set(val)
    return 0


def tramba_batch_sum(src, out, G, B, n, dtype, stream):
    _buf(out, (G, n), dtype).set(_buf(src, (G, B, n), dtype).get().sum(1))
    return 0


def tramba_proj_gather(src, dts, Bs, Cs, plan, K, B, R, N, L, dtype, stream):
    tab = _i32(plan, K * L).reshape(K, L)
    s = _buf(src, (B, K, R + 2 * N, L), dtype).get()
    g = np.stack([s[:, k][:, :, tab[k]] for k in range(K)], axis=1)
    _buf(dts, (B, K, R, L), dtype).set(g[:, :, :R])
    _buf(Bs, (B, K, N, L), dtype).set(g[:, :, R:R + N])
    _buf(Cs, (B, K, N, L), dtype).set(g[:, :, R + N:])
    return 0


def tramba_proj_scatter(gdts, gBs, gCs, gsrc, plan, K, B, R, N, L, dtype, bc_dtype, stream):
    tab = _i32(plan, K * L).reshape(K, L)
    g = np.concatenate([_buf(gdts, (B, K, R, L), dtype).get(), _buf(gBs, (B, K, N, L), bc_dtype).get(),
                        _buf(gCs, (B, K, N, L), bc_dtype).get()], axis=2)
    out = np.zeros((B, K, R + 2 * N, L))
    for k in range(K):
        np.add.at(out[:, k], (slice(None), slice(None), tab[k]), g[:, k])
    _buf(gsrc, (B, K, R + 2 * N, L), dtype).set(out)
    return 0


_ACT = {0: "none", 1: "gelu", 2: "silu"}


def tramba_layernorm2d_fwd(x, w, b, y, mean, rstd, B, C, L, eps, act, in_dtype, out_dtype, stream):
    xv = _buf(x, (B, C, L), in_dtype).get()
    _buf(y, (B, C, L), out_dtype).set(oln.layernorm2d_fwd(xv, _buf(w, (C,)).get(), _get(_buf(b, (C,))), eps, _ACT[act]))
    if mean:
        _buf(mean, (B, L)).set(xv.mean(1))
    if rstd:
        _buf(rstd, (B, L)).set(1.0 / np.sqrt(xv.var(1) + eps))
    return 0


def tramba_layernorm2d_tiles(B, C, L):       # the real one sizes its plan by the device's SM count
    return 1


def tramba_layernorm2d_bwd(gy, x, w, b, mean, rstd, gx, gw_part, gb_part, B, C, L, act, in_dtype, g_dtype, stream):
    xv = _buf(x, (B, C, L), in_dtype).get()
    # eps is not an argument of the backward (the kernel uses the saved rstd): recover it from rstd
    var = xv.var(1)
    eps = float(np.median(1.0 / _buf(rstd, (B, L)).get() ** 2 - var))
    g = oln.layernorm2d_bwd(_buf(gy, (B, C, L), g_dtype).get(), xv, _buf(w, (C,)).get(), _get(_buf(b, (C,))),
                            max(eps, 0.0), _ACT[act])
    _buf(gx, (B, C, L), in_dtype).set(g[0])
    rows = B * tramba_layernorm2d_tiles(B, C, L)
    for ptr, val in ((gw_part, g[1]), (gb_part, g[2])):
        if ptr:
            part = np.zeros((rows, C))
            part[0] = val
            _buf(ptr, (rows, C)).set(part)
    return 0


def tramba_dwconv2d_fwd(x, w, b, y, B, C, H, W, K, act, dtype, stream):
    _buf(y, (B, C, H, W), dtype).set(odw.dwconv_fwd(_buf(x, (B, C, H, W), dtype).get(), _buf(w, (C, 1, K, K)).get(),
                                                    _get(_buf(b, (C,))), _ACT[act]))
    return 0


def tramba_dwconv2d_bwd(x, w, b, dy, dx, dw_part, B, C, H, W, K, act, dtype, stream):
    g = odw.dwconv_bwd(_buf(dy, (B, C, H, W), dtype).get(), _buf(x, (B, C, H, W), dtype).get(),
                       _buf(w, (C, 1, K, K)).get(), _get(_buf(b, (C,))), _ACT[act])
    _buf(dx, (B, C, H, W), dtype).set(g[0])
    part = np.zeros((B, C, K * K + 1))           # per-plane partials the caller sums over B: all in b = 0
    part[0, :, :K * K] = g[1].reshape(C, K * K)
    part[0, :, K * K] = g[2]
    _buf(dw_part, (B, C, K * K + 1)).set(part)
    return 0


def tramba_dct2d_fwd(x, wy, wx, high, low, planes, H, W, stream):
    hi, lo = odct.dct2d_split(_buf(x, (1, planes, H, W)).get())
    _buf(high, (1, planes, H // 2, W // 2)).set(hi)
    _buf(low, (1, planes, H // 2, W // 2)).set(lo)
    return 0


def tramba_dct2d_bwd(ghigh, glow, wy, wx, gx, planes, H, W, stream):
    g = odct.dct2d_split_bwd(_buf(ghigh, (1, planes, H // 2, W // 2)).get(), _buf(glow, (1, planes, H // 2, W // 2)).get(),
                             H, W)
    _buf(gx, (1, planes, H, W)).set(g)
    return 0


_EMULATED = {f.__name__: f for f in (
    tramba_cross_scan, tramba_cross_merge, tramba_sscan_fwd, tramba_sscan_bwd,
    tramba_ss2d_fused_fwd, tramba_ss2d_fused_bwd, tramba_batch_sum,
    tramba_proj_gather, tramba_proj_scatter, tramba_layernorm2d_fwd,
    tramba_layernorm2d_bwd, tramba_layernorm2d_tiles, tramba_dwconv2d_fwd, tramba_dwconv2d_bwd, tramba_dct2d_fwd, tramba_dct2d_bwd)}


class _Proxy:
    def __init__(self, real):
        self._real = real
        self.calls = []

    def __getattr__(self, name):
        fn = _EMULATED.get(name)
        if fn is None:
            return getattr(self._real, name)          # host-only entry points: the real library

        def call(*args):
            self.calls.append(name)
            args = [a.value if isinstance(a, ctypes.c_void_p) else a for a in args]
            return fn(*args)
        return call


class _Stream:
    cuda_stream = 0


class _SimLib:
    """The product's C ABI exported by tools/cpu_sim's library: the REAL entry points (capi.cu) and the REAL kernels
    (every .cu of the library), compiled as host code and executed on the CPU simulator."""

    def __init__(self, sms=4):
        import importlib.util
        import os
        from tramba_b200 import _lib
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        spec = importlib.util.spec_from_file_location("build_sim", os.path.join(root, "tools", "cpu_sim", "build_sim.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        import os as _os
        asan = _os.environ.get("TRAMBA_SIM_ASAN", "") == "1"          # AddressSanitizer build (needs libasan preloaded)
        self._dll = ctypes.CDLL(mod.build(asan=asan))
        for name, (res, args) in _lib.SIGNATURES.items():
            fn = getattr(self._dll, name)
            fn.restype, fn.argtypes = res, args
        self._dll.sim_set_sm_count.restype, self._dll.sim_set_sm_count.argtypes = None, [ctypes.c_int]
        self._dll.sim_set_sm_count(sms)
        self._dll.sim_set_lazy.restype, self._dll.sim_set_lazy.argtypes = None, [ctypes.c_int]
        self._dll.sim_set_lazy(1 if _os.environ.get("TRAMBA_SIM_LAZY", "") == "1" else 0)   # copies land at the wait
        self._dll.sim_set_schedule.restype, self._dll.sim_set_schedule.argtypes = None, [ctypes.c_uint]
        self._dll.sim_set_schedule(int(_os.environ.get("TRAMBA_SIM_SCHEDULE", "0")))       # thread order per sweep
        self.calls = []

    def __getattr__(self, name):
        fn = getattr(self._dll, name)
        if not name.startswith("tramba_"):
            return fn

        def call(*args):
            self.calls.append(name)
            return fn(*args)
        return call


@contextlib.contextmanager
def installed(backend="oracle", sms=4):
    """Within this scope the product's Python host runs on CPU tensors against
      backend="oracle": the compute entry points emulated by the oracle (fast; checks the host logic only), or
      backend="sim":    the product's own C ABI and kernels executed on the CPU simulator (tools/cpu_sim)."""
    from tramba_b200 import _lib
    saved = (_lib.lib, torch.Tensor.cuda, torch.nn.Module.cuda, torch.cuda.current_stream, torch.cuda.device)
    proxy = _Proxy(_lib.lib) if backend == "oracle" else _SimLib(sms)
    _lib.lib = proxy
    torch.Tensor.cuda = lambda self, *a, **k: self.clone()        # like the real one: a new tensor, differentiable
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.Tensor.is_cuda = property(lambda self: True)
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.device = lambda d: contextlib.nullcontext()
    try:
        yield proxy
    finally:
        _lib.lib, torch.Tensor.cuda, torch.nn.Module.cuda, torch.cuda.current_stream, torch.cuda.device = saved
        del torch.Tensor.is_cuda                      # back to the C getter of torch._C.TensorBase
