"""CPU autograd operators backed by the oracle (numpy / C) -- TEST INFRASTRUCTURE ONLY
(imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs).

They have the reference's operator signatures so the same module code can be evaluated on CPU
with the oracle and on the GPU with the product kernels, and the two compared."""
import numpy as np
import torch

from . import scan_orders as so
from . import sscan as osscan
from . import dct as odct

WS = {12: 4, 24: 8, 48: 12, 96: 16, 192: 32}


def _param(fam, H):
    return {"dilation": 4, "window": WS.get(H, 0)}.get(fam, 0)


def make_pair(fam):
    class Scan(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            B, C, H, W = x.shape
            ctx.shape = (B, C, H, W)
            tab = so.order_table(fam, H, _param(fam, H))
            return torch.from_numpy(np.ascontiguousarray(so.cross_scan(x.detach().contiguous().numpy(), tab)))

        @staticmethod
        def backward(ctx, ys):
            B, C, H, W = ctx.shape
            tab = so.order_table(fam, H, _param(fam, H))
            return torch.from_numpy(np.ascontiguousarray(so.cross_merge(ys.contiguous().numpy(), tab, fam))).view(B, C, H, W)

    class Merge(torch.autograd.Function):
        @staticmethod
        def forward(ctx, ys):
            B, K, D, H, W = ys.shape
            ctx.shape = (H, W)
            tab = so.order_table(fam, H, _param(fam, H))
            return torch.from_numpy(np.ascontiguousarray(so.cross_merge(ys.detach().reshape(B, K, D, -1).contiguous().numpy(), tab, fam)))

        @staticmethod
        def backward(ctx, x):
            H, W = ctx.shape
            B, C, L = x.shape
            tab = so.order_table(fam, H, _param(fam, H))
            xs = so.cross_scan(x.contiguous().numpy().reshape(B, C, H, W), tab)
            return torch.from_numpy(np.ascontiguousarray(xs)).view(B, -1, C, H, W)

    return Scan, Merge


class OracleSelectiveScan(torch.autograd.Function):
    """float64 oracle recurrence; returns float32 like oflex=True."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1,
                oflex=True):
        ctx.delta_softplus = delta_softplus
        ctx.save_for_backward(u, delta, A, B, C, D, delta_bias)
        f = lambda t: None if t is None else t.detach().double().numpy()
        out = osscan.selective_scan_fwd(f(u), f(delta), f(A), f(B), f(C), f(D), f(delta_bias), delta_softplus)
        return torch.from_numpy(out).to(torch.float32 if oflex else u.dtype)

    @staticmethod
    def backward(ctx, dout, *args):
        u, delta, A, B, C, D, delta_bias = ctx.saved_tensors
        f = lambda t: None if t is None else t.detach().double().numpy()
        g = osscan.selective_scan_bwd(f(u), f(delta), f(A), f(B), f(C), f(D), f(delta_bias), f(dout),
                                      ctx.delta_softplus)
        ref = (u, delta, A, B, C, D, delta_bias)
        out = tuple(None if (gi is None or r is None) else torch.from_numpy(gi).to(r.dtype) for gi, r in zip(g, ref))
        return out + (None, None, None, None)


class OracleDCT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.hw = x.shape[-2:]
        high, low = odct.dct2d_split(x.detach().double().numpy())
        return torch.from_numpy(np.ascontiguousarray(high)).float(), torch.from_numpy(np.ascontiguousarray(low)).float()

    @staticmethod
    def backward(ctx, gh, gl):
        H, W = ctx.hw
        return torch.from_numpy(odct.dct2d_split_bwd(gh.double().numpy(), gl.double().numpy(), H, W)).float()


FAMILY_OF = {"CrossScan": "std", "CrossScan_Dilation": "dilation", "CrossScan_Window": "window",
             "CrossScan_Line": "line"}


class _OracleDCTModule(torch.nn.Module):
    def forward(self, x):
        return OracleDCT.apply(x.float())


def oracleize(model, f32=False):
    """Swap every product operator inside `model` (tramba_b200.SS2D / DCT2D instances) for the
    CPU oracle operators, in place, so the same module tree can run without a GPU.
    f32=True uses the float32 C oracle (the timed CPU baseline) instead of the float64 truth."""
    import tramba_b200
    from tramba_b200 import dct as _dct

    sel = OracleSelectiveScanF32 if f32 else OracleSelectiveScan
    for m in model.modules():
        if isinstance(m, tramba_b200.SS2D):
            fam = FAMILY_OF[m.scan.__name__]
            m.scan, m.merge = make_pair(fam)
            m.selective_scan = sel
        for name, child in list(m.named_children()):
            if isinstance(child, _dct.DCT2D):
                repl = _OracleDCTModule()
                repl.dct_x, repl.dct_y = child.dct_x, child.dct_y  # keep state_dict keys
                setattr(m, name, repl)
    return model


class OracleSelectiveScanF32(torch.autograd.Function):
    """float32 C oracle (OpenMP): what bench.py times as the CPU baseline."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1,
                oflex=True):
        ctx.delta_softplus = delta_softplus
        ctx.save_for_backward(u, delta, A, B, C, D, delta_bias)
        f = lambda t: None if t is None else t.detach().float().numpy()
        out = osscan.selective_scan_fwd(f(u), f(delta), f(A), f(B), f(C), f(D), f(delta_bias), delta_softplus,
                                        dtype=np.float32)
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, dout, *args):
        u, delta, A, B, C, D, delta_bias = ctx.saved_tensors
        f = lambda t: None if t is None else t.detach().float().numpy()
        g = osscan.selective_scan_bwd(f(u), f(delta), f(A), f(B), f(C), f(D), f(delta_bias), f(dout.contiguous()),
                                      ctx.delta_softplus, dtype=np.float32)
        ref = (u, delta, A, B, C, D, delta_bias)
        out = tuple(None if (gi is None or r is None) else torch.from_numpy(gi).to(r.dtype) for gi, r in zip(g, ref))
        return out + (None, None, None, None)
