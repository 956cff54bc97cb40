"""LayerNorm2d (+ GELU): oracle pinned to the reference class (CPU), CUDA kernel vs oracle (GPU)."""
import numpy as np
import pytest
import torch

from oracle import layernorm2d as oln

CASES = ((2, 8, 6, 6), (1, 64, 12, 12), (2, 256, 5, 7), (1, 130, 3, 11))   # = make_golden_layernorm.py


def _inputs(i):
    B, C, H, W = CASES[i]
    rs = np.random.RandomState(3000 + i)
    x = (rs.standard_normal((B, C, H, W)) * 2.0 + 0.7).astype(np.float32)
    w = (1.0 + 0.3 * rs.standard_normal(C)).astype(np.float32)
    b = (0.2 * rs.standard_normal(C)).astype(np.float32)
    gy = rs.standard_normal((B, C, H, W)).astype(np.float32)
    return x, w, b, gy


@pytest.fixture(scope="module")
def golden_ln():
    import os
    from conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, "layernorm2d.npz"))


@pytest.mark.parametrize("i", range(len(CASES)))
@pytest.mark.parametrize("act", ["none", "gelu"])
def test_oracle_matches_reference_class(golden_ln, i, act):
    x, w, b, gy = _inputs(i)
    y = oln.layernorm2d_fwd(x, w, b, 1e-5, act)
    gx, gw, gb = oln.layernorm2d_bwd(gy, x, w, b, 1e-5, act)
    np.testing.assert_allclose(y, golden_ln[f"ln_{i}_{act}_y"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(gx, golden_ln[f"ln_{i}_{act}_gx"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(gw, golden_ln[f"ln_{i}_{act}_gw"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(gb, golden_ln[f"ln_{i}_{act}_gb"], rtol=1e-4, atol=1e-4)


def test_module_cpu_path_is_the_reference_composition():
    """Without a GPU the harness module runs the reference's own torch ops (no CUDA kernel, no oracle)."""
    from tramba_b200.ss2d import LayerNorm2d
    x, w, b, _ = _inputs(1)
    m = LayerNorm2d(64)
    with torch.no_grad():
        m.weight.copy_(torch.from_numpy(w)), m.bias.copy_(torch.from_numpy(b))
    y = m(torch.from_numpy(x), act="gelu")
    np.testing.assert_allclose(y.detach().numpy(), oln.layernorm2d_fwd(x, w, b, 1e-5, "gelu"), rtol=1e-4, atol=2e-5)


# ------------------------------------------------------------------------------------------ GPU
def _close(a, ref, rtol, atol_scale):
    a = a.detach().double().cpu().numpy()
    tol = rtol * np.abs(ref) + atol_scale * np.abs(ref).max()
    err = np.abs(a - ref)
    assert (err <= tol).all(), f"max err {err.max():.3e} (ref max {np.abs(ref).max():.3e})"


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
@pytest.mark.parametrize("act", ["none", "gelu"])
def test_kernel_vs_reference_golden_fp32(golden_ln, i, act):
    from tramba_b200 import norm
    x, w, b, gy = _inputs(i)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    wt = torch.from_numpy(w).cuda().requires_grad_(True)
    bt = torch.from_numpy(b).cuda().requires_grad_(True)
    y = norm.layer_norm2d(xt, wt, bt, 1e-5, act)
    gx, gw, gb = torch.autograd.grad(y, [xt, wt, bt], torch.from_numpy(gy).cuda())
    # fp32 tolerance of north_star: rtol 1e-3 (+ an absolute floor at 1e-5 of the tensor's scale)
    _close(y, golden_ln[f"ln_{i}_{act}_y"].astype(np.float64), 1e-3, 1e-5)
    _close(gx, golden_ln[f"ln_{i}_{act}_gx"].astype(np.float64), 1e-3, 1e-5)
    _close(gw, golden_ln[f"ln_{i}_{act}_gw"].astype(np.float64), 1e-3, 1e-5)
    _close(gb, golden_ln[f"ln_{i}_{act}_gb"].astype(np.float64), 1e-3, 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("B,C,H,in_dt,out_dt,act", [
    (4, 128, 96, torch.float32, torch.bfloat16, "none"),     # VSSBlock.norm under autocast
    (4, 256, 96, torch.float32, torch.bfloat16, "gelu"),     # SS2D.out_norm + GELU on the merged y
    (4, 1024, 24, torch.bfloat16, torch.bfloat16, "gelu"),
    (4, 2048, 12, torch.float32, torch.float32, "none"),
    (2, 64, 192, torch.bfloat16, torch.float32, "none"),     # patch_embed
    (1, 128, 384, torch.bfloat16, torch.float32, "none"),    # final expand
])
def test_kernel_model_shapes_vs_oracle(B, C, H, in_dt, out_dt, act):
    from tramba_b200 import norm
    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn(B, C, H, H, generator=g) * 1.5 + 0.5).to(in_dt)
    w = 1.0 + 0.2 * torch.randn(C, generator=g)
    b = 0.1 * torch.randn(C, generator=g)
    gy = torch.randn(B, C, H, H, generator=g).to(out_dt)
    xt, wt, bt = (t.cuda().requires_grad_(True) for t in (x, w, b))
    y = norm.layer_norm2d(xt, wt, bt, 1e-5, act, out_dtype=out_dt)
    assert y.dtype == out_dt and y.shape == x.shape
    gx, gw, gb = torch.autograd.grad(y, [xt, wt, bt], gy.cuda())
    assert gx.dtype == in_dt
    sl = (slice(0, 1), slice(None), slice(0, 7), slice(None))    # the oracle on a slab: full C, 7 rows
    xs, gys = x.double().numpy()[sl], gy.double().numpy()[sl]
    ry = oln.layernorm2d_fwd(xs, w.numpy(), b.numpy(), 1e-5, act)
    rgx, _, _ = oln.layernorm2d_bwd(gys, xs, w.numpy(), b.numpy(), 1e-5, act)
    bf = in_dt == torch.bfloat16 or out_dt == torch.bfloat16
    rtol, floor = (1e-2, 4e-3) if bf else (1e-3, 1e-5)          # bf16: rtol 1e-2 + one output ulp at scale
    _close(y[sl], ry, rtol, floor)
    _close(gx[sl], rgx, rtol, floor)
    # dweight / dbias over the FULL tensor vs torch fp64 on the same (rounded) inputs
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    z = torch.nn.functional.layer_norm(xr.permute(0, 2, 3, 1), (C,), wr, br, 1e-5).permute(0, 3, 1, 2)
    if act == "gelu":
        z = torch.nn.functional.gelu(z)
    _, rgw, rgb = torch.autograd.grad(z, [xr, wr, br], gy.double())
    _close(gw, rgw.numpy(), rtol, 1e-3 if bf else 1e-5)
    _close(gb, rgb.numpy(), rtol, 1e-3 if bf else 1e-5)


@pytest.mark.gpu
def test_large_mean_does_not_cancel():
    """Shifted moments: a map with |mean| >> std keeps fp32 accuracy."""
    from tramba_b200 import norm
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 96, 8, 8, generator=g) * 0.01 + 300.0
    w, b = torch.ones(96), torch.zeros(96)
    y = norm.layer_norm2d(x.cuda(), w.cuda(), b.cuda(), 1e-5, "none")
    ref = oln.layernorm2d_fwd(x.numpy(), w.numpy(), b.numpy(), 1e-5, "none")
    _close(y, ref, 1e-3, 1e-3)
