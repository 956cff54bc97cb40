"""Depthwise conv (+ SiLU): oracle pinned to the reference's own torch ops (CPU), CUDA kernel vs oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import dwconv as odw

CASES = ((2, 5, 12, 12, 3, True), (1, 4, 24, 24, 3, False), (2, 3, 10, 14, 5, True), (1, 2, 17, 9, 7, True))


def _inputs(i):
    B, C, H, W, k, has_bias = CASES[i]
    rs = np.random.RandomState(4000 + i)
    x = rs.standard_normal((B, C, H, W)).astype(np.float32)
    w = (0.5 * rs.standard_normal((C, 1, k, k))).astype(np.float32)
    b = (0.3 * rs.standard_normal(C)).astype(np.float32)
    gy = rs.standard_normal((B, C, H, W)).astype(np.float32)
    return x, w, (b if has_bias else None), gy


@pytest.fixture(scope="module")
def golden_dw():
    from conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, "dwconv.npz"))


@pytest.mark.parametrize("i", range(len(CASES)))
@pytest.mark.parametrize("act", ["none", "silu"])
def test_oracle_matches_reference_ops(golden_dw, i, act):
    x, w, b, gy = _inputs(i)
    y = odw.dwconv_fwd(x, w, b, act)
    gx, gw, gb = odw.dwconv_bwd(gy, x, w, b, act)
    np.testing.assert_allclose(y, golden_dw[f"dw_{i}_{act}_y"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gx, golden_dw[f"dw_{i}_{act}_gx"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gw, golden_dw[f"dw_{i}_{act}_gw"], rtol=1e-4, atol=1e-4)
    if b is not None:
        np.testing.assert_allclose(gb, golden_dw[f"dw_{i}_{act}_gb"], rtol=1e-4, atol=1e-4)


def _close(a, ref, rtol, floor):
    a = a.detach().double().cpu().numpy()
    ref = np.asarray(ref, np.float64)
    tol = rtol * np.abs(ref) + floor * np.abs(ref).max()
    err = np.abs(a - ref)
    assert (err <= tol).all(), f"max err {err.max():.3e} (ref max {np.abs(ref).max():.3e})"


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
@pytest.mark.parametrize("act", ["none", "silu"])
def test_kernel_vs_reference_golden_fp32(golden_dw, i, act):
    from tramba_b200 import dwconv
    x, w, b, gy = _inputs(i)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    wt = torch.from_numpy(w).cuda().requires_grad_(True)
    bt = None if b is None else torch.from_numpy(b).cuda().requires_grad_(True)
    assert dwconv.supported(xt, wt)
    y = dwconv.dwconv2d(xt, wt, bt, act)
    grads = torch.autograd.grad(y, [xt, wt] + ([bt] if bt is not None else []), torch.from_numpy(gy).cuda())
    _close(y, golden_dw[f"dw_{i}_{act}_y"], 1e-3, 1e-5)      # fp32: rtol 1e-3 (north_star)
    _close(grads[0], golden_dw[f"dw_{i}_{act}_gx"], 1e-3, 1e-5)
    _close(grads[1], golden_dw[f"dw_{i}_{act}_gw"], 1e-3, 1e-5)
    if bt is not None:
        _close(grads[2], golden_dw[f"dw_{i}_{act}_gb"], 1e-3, 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("B,C,H,k,dt,act", [
    (4, 256, 96, 3, torch.bfloat16, "silu"),     # SS2D.conv2d at enc0 / dec96
    (4, 512, 48, 3, torch.bfloat16, "silu"),
    (4, 1024, 24, 3, torch.float32, "silu"),
    (4, 2048, 12, 3, torch.bfloat16, "silu"),
    (2, 512, 96, 7, torch.bfloat16, "none"),     # DWMSMlp taps
    (2, 64, 48, 5, torch.float32, "none"),
])
def test_kernel_model_shapes_vs_oracle(B, C, H, k, dt, act):
    from tramba_b200 import dwconv
    g = torch.Generator().manual_seed(C + H + k)
    x = torch.randn(B, C, H, H, generator=g).to(dt)
    w = torch.randn(C, 1, k, k, generator=g) / k
    b = 0.2 * torch.randn(C, generator=g)
    gy = torch.randn(B, C, H, H, generator=g).to(dt)
    xt, wt, bt = (t.cuda().requires_grad_(True) for t in (x, w, b))
    y = dwconv.dwconv2d(xt, wt, bt, act)
    assert y.dtype == dt and y.shape == x.shape
    gx, gw, gb = torch.autograd.grad(y, [xt, wt, bt], gy.cuda())
    cs = slice(0, 6)                                           # the oracle on a few channels, all of B
    ry = odw.dwconv_fwd(x[:, cs].double().numpy(), w[cs].numpy(), b[cs].numpy(), act)
    rgx, rgw, rgb = odw.dwconv_bwd(gy[:, cs].double().numpy(), x[:, cs].double().numpy(), w[cs].numpy(),
                                   b[cs].numpy(), act)
    bf = dt == torch.bfloat16
    rtol, floor = (1e-2, 4e-3) if bf else (1e-3, 1e-5)
    _close(y[:, cs], ry, rtol, floor)
    _close(gx[:, cs], rgx, rtol, floor)
    _close(gw[cs], rgw, rtol, 1e-3 if bf else 1e-5)
    _close(gb[cs], rgb, rtol, 1e-3 if bf else 1e-5)
