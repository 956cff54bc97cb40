"""GPU parity: cross-scan / cross-merge through the C-ABI vs the reference's golden outputs
(small sizes) and vs the oracle (full sizes).  Integer/copy work: bit-exact."""
import numpy as np
import pytest
import torch

import tramba_b200
from tramba_b200 import _lib, ops
from oracle import scan_orders as so

pytestmark = pytest.mark.gpu
WS = {12: 4, 24: 8, 48: 12, 96: 16, 192: 32}
PAIRS = {
    "std": (tramba_b200.CrossScan, tramba_b200.CrossMerge, 4),
    "dilation": (tramba_b200.CrossScan_Dilation, tramba_b200.CrossMerge_Dilation, 4),
    "window": (tramba_b200.CrossScan_Window, tramba_b200.CrossMerge_Window, 4),
    "line": (tramba_b200.CrossScan_Line, tramba_b200.CrossMerge_Line, 8),
}
ORDER = {"std": _lib.ORDER_STD, "dilation": _lib.ORDER_DILATION, "window": _lib.ORDER_WINDOW,
         "line": _lib.ORDER_LINE}


def _param(fam, H):
    return {"std": 0, "dilation": 4, "window": WS.get(H, 0), "line": 0}[fam]


@pytest.mark.parametrize("fam", list(PAIRS))
@pytest.mark.parametrize("H", [12, 24])
def test_against_reference_golden(golden_scanmerge, fam, H):
    """Same seeded inputs as tests/golden/make_golden.py; outputs of the real reference."""
    g = golden_scanmerge
    S, M, K = PAIRS[fam]
    rs = np.random.RandomState(1000 + H)
    x = rs.standard_normal((2, 3, H, H)).astype(np.float32)
    ys = rs.standard_normal((2, K, 3, H, H)).astype(np.float32)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    xs = S.apply(xt)
    assert np.array_equal(xs.detach().cpu().numpy(), g[f"{fam}_{H}_scan"])
    (gx,) = torch.autograd.grad(xs, xt, torch.from_numpy(ys).cuda().view(2, K, 3, H * H))
    assert np.array_equal(gx.cpu().numpy(), g[f"{fam}_{H}_scan_bwd"])
    yt = torch.from_numpy(ys).cuda().requires_grad_(True)
    y = M.apply(yt)
    assert np.array_equal(y.detach().cpu().numpy(), g[f"{fam}_{H}_merge"])
    (gys,) = torch.autograd.grad(y, yt, torch.from_numpy(x).cuda().view(2, 3, H * H))
    assert np.array_equal(gys.cpu().numpy(), g[f"{fam}_{H}_merge_bwd"])


@pytest.mark.parametrize("fam", list(PAIRS))
@pytest.mark.parametrize("H,B,C", [(96, 2, 5), (48, 1, 7), (192, 1, 2)])
def test_full_size_vs_oracle(fam, H, B, C):
    tab = so.order_table(fam, H, _param(fam, H))
    K = tab.shape[0]
    rs = np.random.RandomState(H + B)
    x = rs.standard_normal((B, C, H, H)).astype(np.float32)
    ys = rs.standard_normal((B, K, C, H * H)).astype(np.float32)
    xs = tramba_b200.cross_scan(torch.from_numpy(x).cuda(), ORDER[fam], _param(fam, H))
    assert np.array_equal(xs.cpu().numpy(), so.cross_scan(x, tab))
    y = tramba_b200.cross_merge(torch.from_numpy(ys).cuda(), ORDER[fam], H, H, _param(fam, H))
    assert np.array_equal(y.cpu().numpy(), so.cross_merge(ys, tab, fam))


@pytest.mark.parametrize("fam", list(PAIRS))
def test_bf16(fam):
    H, B, C = 24, 2, 8
    tab = so.order_table(fam, H, _param(fam, H))
    K = tab.shape[0]
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, H, H, generator=g).bfloat16()
    ys = torch.randn(B, K, C, H * H, generator=g).bfloat16()
    xs = tramba_b200.cross_scan(x.cuda(), ORDER[fam], _param(fam, H))
    assert xs.dtype == torch.bfloat16
    assert np.array_equal(xs.float().cpu().numpy(), so.cross_scan(x.float().numpy(), tab))
    y = tramba_b200.cross_merge(ys.cuda(), ORDER[fam], H, H, _param(fam, H))
    ref = torch.from_numpy(so.cross_merge(ys.float().numpy(), tab, fam)).bfloat16()   # fp32 sums, one rounding
    assert torch.equal(y.cpu(), ref)


@pytest.mark.parametrize("H", [6, 7, 14])
def test_line_small_and_odd_sizes(H):
    """L % 4 != 0 exercises the scalar paths; H = 7 is an odd-size table the reference also builds."""
    tab = np.concatenate([so.std_indices(H, H), np.stack(so.line_indices(H, H))])   # order_table() insists on even H
    rs = np.random.RandomState(H)
    x = rs.standard_normal((2, 3, H, H)).astype(np.float32)
    ys = rs.standard_normal((2, 8, 3, H * H)).astype(np.float32)
    ops.line_sizes.append(H)
    xs = tramba_b200.cross_scan(torch.from_numpy(x).cuda(), _lib.ORDER_LINE)
    assert np.array_equal(xs.cpu().numpy(), so.cross_scan(x, tab))
    y = tramba_b200.cross_merge(torch.from_numpy(ys).cuda(), _lib.ORDER_LINE, H, H)
    assert np.array_equal(y.cpu().numpy(), so.cross_merge(ys, tab, "line"))


@pytest.mark.parametrize("fam", ["std", "dilation", "window"])
def test_roundtrip_4x_full_size(fam):
    """Dilation.py:115 / Window.py:105 invariant at BASELINE size (B=4, D=256, 96x96)."""
    x = torch.randn(4, 256, 96, 96, device="cuda")
    xs = tramba_b200.cross_scan(x, ORDER[fam], _param(fam, 96))
    y = tramba_b200.cross_merge(xs, ORDER[fam], 96, 96, _param(fam, 96))
    assert torch.equal(y.view_as(x), 4 * x)


def test_line_roundtrip_multiplicity_and_adjoint_full_size():
    H = 96
    x = torch.randn(4, 256, H, H, device="cuda")
    xs = tramba_b200.cross_scan(x, _lib.ORDER_LINE)
    tab = torch.from_numpy(so.order_table("line", H)).cuda()
    mult = torch.zeros(H * H, device="cuda").index_add_(0, tab.flatten(), torch.ones(tab.numel(), device="cuda"))
    ones = tramba_b200.cross_merge(torch.ones(1, 8, 1, H * H, device="cuda"), _lib.ORDER_LINE, H, H)
    assert torch.equal(ones.flatten(), mult)
    g = torch.randn_like(xs)
    lhs = (xs.double() * g.double()).sum()
    rhs = (x.double().flatten(2) * tramba_b200.cross_merge(g, _lib.ORDER_LINE, H, H).double()).sum()
    assert abs(lhs - rhs) <= 1e-6 * abs(lhs) + 1e-3


def test_keyerror_like_reference():
    with pytest.raises(KeyError):
        tramba_b200.CrossScan_Window.apply(torch.zeros(1, 1, 20, 20, device="cuda"))
