"""float64 CPU statement of the fused SS2D core (test infrastructure): gather the image-ordered x
into scan order with the oracle's order table, run the oracle selective scan on it with the
scan-ordered delta / B / C, scatter-add every token back to its pixel (Models/vmamba.py:231-259 as
one function)."""
import numpy as np
import torch

from oracle import scan_orders as so
from oracle import sscan as osscan

WS = {12: 4, 24: 8, 48: 12, 96: 16, 192: 32}


def param_of(fam, H):
    return {"std": 0, "dilation": 4, "window": WS.get(H, 0), "line": 0}[fam]


class _Sel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, delta, A_, B_, C_, D_, b_):
        ctx.save_for_backward(u, delta, A_, B_, C_, D_, b_)
        f = lambda t: t.detach().numpy()
        return torch.from_numpy(osscan.selective_scan_fwd(f(u), f(delta), f(A_), f(B_), f(C_), f(D_), f(b_), True))

    @staticmethod
    def backward(ctx, g):
        f = lambda t: t.detach().numpy()
        gs = osscan.selective_scan_bwd(*[f(t) for t in ctx.saved_tensors], f(g.contiguous()), True)
        return tuple(torch.from_numpy(np.ascontiguousarray(v)) for v in gs)


def fused_core_oracle(fam, H, tensors, dy=None):
    """tensors = [x (B,D,H,H) image order, dts (B,K,D,L), Bs (B,K,L), Cs (B,K,L) scan order, A, Ds, bias (K*D)].
    Returns (y (B,D,L) float64, [grads wrt the 7 tensors] or None)."""
    x, dts, Bs, Cs, A, Ds, bias = [t.detach().double().requires_grad_(True) for t in tensors]
    B, D = x.shape[:2]
    K = dts.shape[1]
    L = H * H
    tab = torch.from_numpy(so.order_table(fam, H, param_of(fam, H)).astype(np.int64))  # (K, L)
    t4 = tab.view(1, K, 1, L).expand(B, K, D, L)
    xs = x.view(B, 1, D, L).expand(B, K, D, L).gather(3, t4)
    dsc = dts
    Bsc = Bs.view(B, K, 1, L)
    Csc = Cs.view(B, K, 1, L)
    ys = _Sel.apply(xs.reshape(B, K * D, L), dsc.reshape(B, K * D, L), A.view(K * D, 1), Bsc, Csc, Ds, bias)
    # merge = transpose of the gather: every token is added to its pixel
    y = torch.zeros(B, D, L, dtype=torch.float64).scatter_add(
        2, t4.permute(0, 2, 1, 3).reshape(B, D, K * L), ys.view(B, K, D, L).permute(0, 2, 1, 3).reshape(B, D, K * L))
    if dy is None:
        return y.detach(), None
    y.backward(dy.double())
    return y.detach(), [t.grad for t in (x, dts, Bs, Cs, A, Ds, bias)]
