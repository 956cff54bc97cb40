"""GPU parity of the fused SS2D core (tramba_ss2d_fused_fwd / _bwd through the C-ABI) against the
oracle composition CrossScan -> selective scan -> CrossMerge (Models/vmamba.py:231-259) evaluated
on CPU in float64, on identical image-ordered operands.

Tolerances (BASELINE.json): outputs and gradients rtol 1e-3 in fp32, 1e-2 in bf16 (with an
absolute floor relative to the tensor's max); scan/merge INDEXING is checked bit-exactly with
integer-valued inputs whose sums are exact in fp32."""
import numpy as np
import pytest
import torch

import tramba_b200
from tramba_b200 import _lib, fused
from _fused_oracle import fused_core_oracle as _oracle, param_of as _param
from oracle import scan_orders as so

pytestmark = pytest.mark.gpu

FAM = {"std": (_lib.ORDER_STD, 4), "dilation": (_lib.ORDER_DILATION, 4), "window": (_lib.ORDER_WINDOW, 4),
       "line": (_lib.ORDER_LINE, 8)}
def _inputs(fam, H, D, B, seed, dtype=torch.float32):
    order, K = FAM[fam]
    L = H * H
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, D, H, H, generator=g)
    dts = torch.randn(B, K, D, L, generator=g)
    Bs, Cs = torch.randn(B, K, L, generator=g), torch.randn(B, K, L, generator=g)
    A = -(0.2 + torch.rand(K * D, generator=g))
    Ds = 1 + 0.1 * torch.randn(K * D, generator=g)
    bias = torch.randn(K * D, generator=g) - 2.0
    ts = [x, dts, Bs, Cs]
    if dtype != torch.float32:  # round once so CPU and GPU start from identical values
        ts = [t.to(dtype).float() for t in ts]
    return order, K, ts + [A, Ds, bias]


def _close(a, b, rtol, what, floor=0.05):
    """|a - b| <= rtol * |b| + floor * rtol * max|b| -- the same absolute floor as test_gpu_sscan.py."""
    a = a.detach().double().cpu().numpy()
    b = b.detach().double().cpu().numpy()
    tol = rtol * np.abs(b) + floor * rtol * np.abs(b).max()
    err = np.abs(a - b)
    assert (err <= tol).all(), f"{what}: max err {err.max():.3e} (ref max {np.abs(b).max():.3e}) at {np.argmax(err - tol)}"


CASES = [("std", 12, 64), ("std", 24, 32), ("std", 48, 8), ("std", 96, 2), ("dilation", 24, 16), ("dilation", 96, 1),
         ("window", 24, 16), ("window", 48, 4), ("window", 96, 1), ("line", 12, 64), ("line", 24, 16), ("line", 48, 4),
         ("line", 96, 1)]


@pytest.mark.parametrize("fam,H,D", CASES)
def test_fused_fwd_bwd_fp32_vs_oracle(fam, H, D):
    B = 2
    order, K, tensors = _inputs(fam, H, D, B, seed=H + D)
    assert fused.supported(order, K, D, H, H, 1, torch.float32, need_bwd=False)
    dev = [t.cuda().requires_grad_(True) for t in tensors]
    y = fused.ss2d_core_fused(*dev, order, _param(fam, H))
    dy = torch.randn(B, D, H * H, generator=torch.Generator().manual_seed(5))
    yr, gr = _oracle(fam, H, tensors, dy)
    _close(y, yr, 1e-3, "y")
    assert fused.supported(order, K, D, H, H, 1, torch.float32, need_bwd=True)    # 96 x 96 in fp32 included
    grads = torch.autograd.grad(y, dev, dy.cuda())
    for name, g, r in zip("dx ddts dBs dCs dA dDs dbias".split(), grads, gr):
        _close(g, r, 1e-3, name)


@pytest.mark.parametrize("fam,H,D", [("std", 96, 2), ("std", 24, 32), ("line", 96, 1), ("line", 24, 16),
                                     ("window", 48, 4), ("dilation", 96, 1)])
def test_fused_fwd_bwd_bf16_vs_oracle(fam, H, D):
    B = 2
    order, K, tensors = _inputs(fam, H, D, B, seed=H + D + 1, dtype=torch.bfloat16)
    dev = [t.cuda() for t in tensors]
    dev = [t.to(torch.bfloat16).requires_grad_(True) for t in dev[:4]] + [t.requires_grad_(True) for t in dev[4:]]
    assert fused.supported(order, K, D, H, H, 1, torch.bfloat16, need_bwd=True)
    y = fused.ss2d_core_fused(*dev, order, _param(fam, H))
    assert y.dtype == torch.float32
    dy = torch.randn(B, D, H * H, generator=torch.Generator().manual_seed(6))
    grads = torch.autograd.grad(y, dev, dy.cuda())
    yr, gr = _oracle(fam, H, tensors, dy)
    _close(y, yr, 1e-3, "y")        # bf16 inputs are exact in both; internals fp32 vs fp64
    for name, g, r in zip("dx ddts dBs dCs dA dDs dbias".split(), grads, gr):
        # ddelta / dB / dC / dx are rounded to bf16 once on the way out (dx is accumulated over K in fp32)
        _close(g, r, 1e-2, name)


@pytest.mark.parametrize("fam,H,D", [("std", 24, 16), ("dilation", 48, 4), ("window", 96, 1), ("line", 24, 16),
                                     ("line", 96, 1)])
def test_fused_indexing_bit_exact(fam, H, D):
    """B = 0, C = 1, Ds = small integers: y[p] = sum_k Ds_k * (#tokens of k at p) * x[p] -- integer
    arithmetic, exact in fp32 -- equals oracle merge(scan(x) * Ds) bit for bit; and with x = pixel id
    the gather side is pinned too."""
    order, K = FAM[fam]
    B, L = 2, H * H
    rs = np.random.RandomState(3)
    x = rs.randint(-8, 9, size=(B, D, H, H)).astype(np.float32)
    x[0, 0] = np.arange(L, dtype=np.float32).reshape(H, H) % 251
    Ds = rs.randint(1, 4, size=(K * D,)).astype(np.float32)
    dts = rs.standard_normal((B, K, D, L)).astype(np.float32)
    zeros, ones = np.zeros((B, K, L), np.float32), np.ones((B, K, L), np.float32)
    A = -np.ones(K * D, np.float32)
    t = lambda a: torch.from_numpy(a).cuda()
    y = fused.ss2d_core_fused(t(x), t(dts), t(zeros), t(ones), t(A), t(Ds), None, order, _param(fam, H))
    tab = so.order_table(fam, H, _param(fam, H))
    xs = so.cross_scan(x, tab) * Ds.reshape(1, K, D, 1)
    ref = so.cross_merge(np.ascontiguousarray(xs), tab, fam)
    assert np.array_equal(y.cpu().numpy(), ref)


def test_ss2d_module_fused_matches_unfused():
    # the two routes evaluate x_proj / dt_proj differently (one GEMM on the image-ordered map vs grouped conv1d on
    # xs): keep cuDNN / cuBLAS out of TF32 so that the comparison is about OUR kernels
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    for scan, merge, K, H in ((tramba_b200.CrossScan, tramba_b200.CrossMerge, 4, 24),
                              (tramba_b200.CrossScan_Line, tramba_b200.CrossMerge_Line, 8, 24),
                              (tramba_b200.CrossScan_Window, tramba_b200.CrossMerge_Window, 4, 48)):
        m = tramba_b200.SS2D(d_model=32, d_state=1, channel_first=True, scan=scan, merge=merge, k_group=K).cuda()
        x = torch.randn(2, 32, H, H, device="cuda")
        outs = []
        for f in (True, False):
            m.fused = f
            m.zero_grad()
            xi = x.clone().requires_grad_(True)
            n0 = _lib.launch_count()
            y = m(xi)
            y.square().sum().backward()
            outs.append((y.detach(), xi.grad, [p.grad.clone() for p in m.parameters()], _lib.launch_count() - n0))
        (yf, gxf, gpf, nf), (yu, gxu, gpu, nu) = outs
        assert nf <= nu                                 # fused: 1 fwd + 1 bwd core launch (+ 2 weight-gradient batch sums) instead of 6
        _close(yf, yu, 1e-3, "y")
        _close(gxf, gxu, 1e-3, "dx")
        for (n, _), a, b in zip(m.named_parameters(), gpf, gpu):
            _close(a, b, 2e-3, n)
