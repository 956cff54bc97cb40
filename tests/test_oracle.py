"""Pins the CPU oracle against the reference's golden fixtures (tests/golden/*.npz, generated
from the real reference by tests/golden/make_golden.py) and against independent statements."""
import numpy as np
import pytest

from oracle import scan_orders as so
from oracle import dct as odct
from oracle import sscan as osscan

WS = {12: 4, 24: 8, 48: 12, 96: 16, 192: 32}


@pytest.mark.parametrize("H", [12, 24, 48, 96, 192])
def test_dilation_window_tables_bit_exact(golden_orders, H):
    assert np.array_equal(so.dilation_indices(H, H, 4), golden_orders[f"dilation_{H}_4"])
    assert np.array_equal(so.window_indices(H, H, WS[H]), golden_orders[f"window_{H}_{WS[H]}"])


def test_reference_main_block_cases(golden_orders):
    # Dilation.py:99-115 (H=4, rate 3) and Window.py:89-105 (H=8, ws=4)
    assert np.array_equal(so.dilation_indices(4, 4, 3), golden_orders["dilation_4_3"])
    assert np.array_equal(so.window_indices(8, 8, 4), golden_orders["window_8_4"])


@pytest.mark.parametrize("H", [6, 7, 12, 14, 24, 28, 48, 56, 96, 192])
def test_line_tables_bit_exact(golden_orders, H):
    got = np.stack(so.line_indices(H, H))
    assert np.array_equal(got, golden_orders[f"line_{H}"])


@pytest.mark.parametrize("H", [12, 96])
def test_line_multiplicity_facts(H):
    # SURVEY A.4 [probe]: 75 % coverage per list, centre multiplicity H/2, all pixels hit >= 2x
    ln = so.line_indices(H, H)
    tot = np.zeros(H * H, int)
    for a in ln:
        cnt = np.bincount(a, minlength=H * H)
        assert (cnt > 0).sum() == 3 * H * H // 4
        assert cnt.max() == H // 2
        tot += cnt
    assert tot.min() >= 2


@pytest.mark.parametrize("fam,param", [("std", 0), ("dilation", 4), ("window", None), ("line", 0)])
@pytest.mark.parametrize("H", [12, 24])
def test_scan_merge_match_reference(golden_scanmerge, fam, param, H):
    g = golden_scanmerge
    if fam == "window":
        param = WS[H]
    tab = so.order_table(fam, H, param)
    K = tab.shape[0]
    rs = np.random.RandomState(1000 + H)
    x = rs.standard_normal((2, 3, H, H)).astype(np.float32)
    ys = rs.standard_normal((2, K, 3, H, H)).astype(np.float32).reshape(2, K, 3, H * H)
    assert np.array_equal(so.cross_scan(x, tab), g[f"{fam}_{H}_scan"])
    assert np.array_equal(so.cross_scan(x, tab).reshape(2, K, 3, H, H), g[f"{fam}_{H}_merge_bwd"])
    # merges: fp32, same association order as the reference's CPU evaluation => bit-exact
    assert np.array_equal(so.cross_merge(ys, tab, fam), g[f"{fam}_{H}_merge"])
    assert np.array_equal(so.cross_merge(ys, tab, fam).reshape(2, 3, H, H), g[f"{fam}_{H}_scan_bwd"])


@pytest.mark.parametrize("fam,param", [("std", 0), ("dilation", 4), ("window", 16)])
def test_roundtrip_4x(fam, param):
    # Dilation.py:115 / Window.py:105: merge(scan(x)) == 4*x for permutation orders
    H = 96
    tab = so.order_table(fam, H, param)
    x = np.random.RandomState(0).standard_normal((1, 2, H, H)).astype(np.float32)
    y = so.cross_merge(so.cross_scan(x, tab), tab, fam)
    assert np.array_equal(y, 4 * x.reshape(1, 2, -1))


def test_line_roundtrip_multiplicity():
    H = 12
    tab = so.order_table("line", H)
    x = np.ones((1, 1, H, H), np.float32)
    y = so.cross_merge(so.cross_scan(x, tab), tab, "line")
    assert set(np.unique(y).tolist()) == {6.0, 8.0, 12.0, 28.0}  # SURVEY A.4 [probe]


@pytest.mark.parametrize("B,C,H", [(2, 3, 24), (1, 2, 48), (1, 1, 96), (2, 2, 12)])
def test_dct_matches_reference(golden_dct, B, C, H):
    g = golden_dct
    assert np.array_equal(odct.dct_basis(H), g[f"dct_{H}_wx"])
    assert np.array_equal(odct.dct_basis(H), g[f"dct_{H}_wy"])
    x = np.random.RandomState(2000 + H).standard_normal((B, C, H, H)).astype(np.float32)
    high, low = odct.dct2d_split(x)
    np.testing.assert_allclose(high, g[f"dct_{B}_{C}_{H}_high"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(low, g[f"dct_{B}_{C}_{H}_low"], rtol=0, atol=2e-5)


def test_dct_matches_scipy():
    import scipy.fft
    x = np.random.RandomState(5).standard_normal((2, 2, 48, 48))
    y = odct.dct2d_full(x)
    np.testing.assert_allclose(y, scipy.fft.dctn(x, type=2, norm="ortho", axes=(2, 3)), atol=1e-6)


def test_dct_bwd_is_adjoint():
    rs = np.random.RandomState(7)
    x = rs.standard_normal((1, 2, 24, 24))
    gh, gl = rs.standard_normal((1, 2, 12, 12)), rs.standard_normal((1, 2, 12, 12))
    high, low = odct.dct2d_split(x)
    lhs = (high * gh).sum() + (low * gl).sum()
    rhs = (x * odct.dct2d_split_bwd(gh, gl, 24, 24)).sum()
    assert abs(lhs - rhs) < 1e-9 * max(1, abs(lhs))


def _rand_sscan(rs, b, K, Dg, L, N):
    dim = K * Dg
    u = rs.standard_normal((b, dim, L))
    delta = rs.standard_normal((b, dim, L))
    A = -np.exp(rs.standard_normal((dim, N)) * 0.5)
    Bm = rs.standard_normal((b, K, N, L))
    Cm = rs.standard_normal((b, K, N, L))
    D = rs.standard_normal(dim)
    bias = rs.standard_normal(dim) - 2.0
    return u, delta, A, Bm, Cm, D, bias


@pytest.mark.parametrize("N", [1, 4, 16])
def test_sscan_c_vs_numpy(N):
    rs = np.random.RandomState(11)
    args = _rand_sscan(rs, 2, 2, 3, 37, N)
    a = osscan.selective_scan_fwd(*args)
    b = osscan.selective_scan_fwd_numpy(*args)
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)
    # softplus threshold (x > 20 -> identity) and no-softplus branch
    args2 = list(args)
    args2[1] = args2[1] * 30
    np.testing.assert_allclose(osscan.selective_scan_fwd(*args2),
                               osscan.selective_scan_fwd_numpy(*args2), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(osscan.selective_scan_fwd(*args, delta_softplus=False),
                               osscan.selective_scan_fwd_numpy(*args, delta_softplus=False),
                               rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("N", [1, 3])
def test_sscan_bwd_vs_torch_autograd(N):
    """Analytic backward of the oracle vs autograd through an independent torch statement of
    selective_scan_ref in fp64 (the strongest pin available: the recurrence is not in the
    reference tree)."""
    import torch
    rs = np.random.RandomState(13)
    b, K, Dg, L = 2, 2, 3, 29
    u, delta, A, Bm, Cm, D, bias = _rand_sscan(rs, b, K, Dg, L, N)
    dout = rs.standard_normal(u.shape)
    tt = [torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (u, delta, A, Bm, Cm, D, bias)]
    tu, td, tA, tB, tC, tD, tb = tt
    dl = torch.nn.functional.softplus(td + tb[None, :, None])
    Bf = tB.repeat_interleave(Dg, dim=1)
    Cf = tC.repeat_interleave(Dg, dim=1)
    h = torch.zeros(b, K * Dg, N, dtype=torch.float64)
    ys = []
    for t in range(L):
        h = torch.exp(dl[:, :, t, None] * tA[None]) * h + dl[:, :, t, None] * Bf[:, :, :, t] * tu[:, :, t, None]
        ys.append((Cf[:, :, :, t] * h).sum(-1))
    out = torch.stack(ys, -1) + tu * tD[None, :, None]
    np.testing.assert_allclose(out.detach().numpy(),
                               osscan.selective_scan_fwd(u, delta, A, Bm, Cm, D, bias), rtol=1e-12, atol=1e-12)
    grads = torch.autograd.grad(out, tt, torch.tensor(dout))
    got = osscan.selective_scan_bwd(u, delta, A, Bm, Cm, D, bias, dout)
    for name, g, o in zip("du ddelta dA dB dC dD dbias".split(), grads, got):
        np.testing.assert_allclose(o, g.numpy(), rtol=1e-10, atol=1e-11, err_msg=name)


def test_sscan_f32_close_to_f64():
    rs = np.random.RandomState(17)
    args = _rand_sscan(rs, 1, 4, 8, 144, 1)
    a = osscan.selective_scan_fwd(*args, dtype=np.float32)
    b = osscan.selective_scan_fwd(*args, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-4)
