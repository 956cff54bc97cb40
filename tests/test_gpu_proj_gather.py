"""GPU parity of tramba_proj_gather / tramba_proj_scatter (the projected-channel gather of the fused SS2D
host path) against the oracle's order tables: the gather is bit-exact; the adjoint is checked bit-exactly
with integer-valued gradients (exact fp32 sums) and against autograd of torch.gather otherwise."""
import numpy as np
import pytest
import torch

from tramba_b200 import _lib, fused
from _fused_oracle import param_of
from oracle import scan_orders as so

pytestmark = pytest.mark.gpu

FAM = {"std": (_lib.ORDER_STD, 4), "dilation": (_lib.ORDER_DILATION, 4), "window": (_lib.ORDER_WINDOW, 4),
       "line": (_lib.ORDER_LINE, 8)}


@pytest.mark.parametrize("fam,H", [("std", 12), ("std", 96), ("dilation", 24), ("window", 48), ("line", 24), ("line", 96)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gather_and_adjoint(fam, H, dtype):
    order, K = FAM[fam]
    B, R, N, L = 2, 5, 1, H * H
    param = param_of(fam, H)
    tab = torch.from_numpy(so.order_table(fam, H, param).astype(np.int64))            # (K, L)
    g = torch.Generator().manual_seed(H)
    src = torch.randint(-8, 9, (B, K, R + 2 * N, L), generator=g).to(dtype)            # exactly representable
    ref = src.gather(3, tab.view(1, K, 1, L).expand_as(src))
    x = src.cuda().requires_grad_(True)
    dts, Bs, Cs = fused.proj_gather(x, order, param, H, R, N)
    assert dts.is_contiguous() and Bs.is_contiguous() and Cs.is_contiguous()
    assert torch.equal(dts.cpu(), ref[:, :, :R]) and torch.equal(Bs.cpu(), ref[:, :, R:R + N])
    assert torch.equal(Cs.cpu(), ref[:, :, R + N:])
    # adjoint with integer-valued gradients: sums are exact, so the result is independent of order
    gd = torch.randint(-4, 5, dts.shape, generator=g).to(dtype)
    gb = torch.randint(-4, 5, Bs.shape, generator=g).float()
    gc = torch.randint(-4, 5, Cs.shape, generator=g).float()
    (gx,) = torch.autograd.grad([dts, Bs, Cs], x, [gd.cuda(), gb.cuda().to(dtype), gc.cuda().to(dtype)])
    gall = torch.cat([gd.float(), gb, gc], 2)
    refg = torch.zeros(B, K, R + 2 * N, L).scatter_add_(3, tab.view(1, K, 1, L).expand_as(gall), gall)
    if dtype == torch.float32 or refg.abs().max() <= 256:                               # exact in bf16 too
        assert torch.equal(gx.float().cpu(), refg.to(dtype).float())
    else:
        torch.testing.assert_close(gx.float().cpu(), refg, rtol=1e-2, atol=1.0)


def test_adjoint_is_the_transpose():
    """<gather(x), g> == <x, scatter(g)> on random fp32 data (helix: duplicates and unvisited pixels)."""
    order, K = FAM["line"]
    B, R, N, H = 1, 3, 1, 24
    L = H * H
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, K, R + 2 * N, L, generator=g).cuda().requires_grad_(True)
    outs = fused.proj_gather(x, order, 0, H, R, N)
    gs = [torch.randn(o.shape, generator=g).cuda() for o in outs]
    (gx,) = torch.autograd.grad(outs, x, gs)
    lhs = sum((o.double() * gi.double()).sum() for o, gi in zip(outs, gs))
    rhs = (x.double() * gx.double()).sum()
    assert abs(lhs.item() - rhs.item()) <= 1e-5 * abs(lhs.item()) + 1e-6


@pytest.mark.parametrize("shape,groups,dtype", [
    ((4, 192, 1024), 0, torch.bfloat16),        # x_proj weight-gradient batch sum (vector kernel)
    ((4, 8, 6, 37), 0, torch.float32),          # odd column count: scalar path
    ((576, 256), 0, torch.float32),             # per-tile partials: tall kernel
    ((2, 576, 256), 2, torch.float32),          # LayerNorm2d: dweight | dbias in one launch
    ((2, 100, 48), 2, torch.float32),           # tall, columns not a multiple of 32
    ((130, 45), 0, torch.bfloat16),
    ((2, 5, 64), 2, torch.float32),             # groups on the vector kernel
])
def test_batch_sum(shape, groups, dtype):
    """tramba_batch_sum: out = sum over the batch axis with fp32 accumulation, same dtype out (both kernels, groups)."""
    g = torch.Generator().manual_seed(sum(shape))
    t = torch.randn(shape, generator=g).to(dtype)
    ref = t.double().sum(1 if groups else 0)
    out = fused._batch_sum(t.cuda(), groups=groups)
    assert out.dtype == dtype and tuple(out.shape) == tuple(ref.shape)
    tol = 1e-5 if dtype == torch.float32 else 2.0 ** -8
    err = (out.double().cpu() - ref).abs().max() / ref.abs().max()
    assert err < tol, float(err)
    ti = torch.randint(-4, 5, shape, generator=g).to(dtype)                              # exact sums
    assert torch.equal(fused._batch_sum(ti.cuda(), groups=groups).cpu().double(), ti.double().sum(1 if groups else 0))
