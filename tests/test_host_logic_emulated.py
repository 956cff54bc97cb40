"""Host logic of the product (autograd wiring, marshalling, routing) on CPU: tramba_b200's Python host runs
unmodified against tests/_abi_emulator.py -- the compute entry points of the C-ABI emulated by the oracle behind the
same raw pointers -- and is compared with the reference-shaped composition of the oracle operators.  No kernel runs
here (the `-m gpu` tests own kernel parity); what is checked is everything between the reference's operator
interface and the C-ABI call."""
import copy

import numpy as np
import pytest
import torch

import tramba_b200
from tramba_b200 import _lib
from _abi_emulator import installed
from _oracle_ops import OracleSelectiveScan, make_pair
from oracle import scan_orders as so

CASES = [
    ("std", "CrossScan", "CrossMerge", 4, 1, 12),
    ("line", "CrossScan_Line", "CrossMerge_Line", 8, 1, 12),
    ("window", "CrossScan_Window", "CrossMerge_Window", 4, 1, 24),
    ("dilation", "CrossScan_Dilation", "CrossMerge_Dilation", 4, 2, 12),      # N = 2: op-by-op path
    ("spiral", "CrossScan_Spiral", "CrossMerge_Spiral", 8, 1, 12),            # ablation order: never fused
]


@pytest.mark.parametrize("fam,scan,merge,K,N,H", CASES)
def test_ss2d_module_host_logic(fam, scan, merge, K, N, H):
    torch.manual_seed(0)
    m = tramba_b200.SS2D(d_model=32, d_state=N, ssm_ratio=2.0, dt_rank="auto", d_conv=3, conv_bias=False,
                         channel_first=True, disable_z=True, scan=getattr(tramba_b200, scan),
                         merge=getattr(tramba_b200, merge), k_group=K)
    ref = copy.deepcopy(m)
    ref.scan, ref.merge = make_pair(fam)
    ref.selective_scan = OracleSelectiveScan
    ref.fused_dwconv = False
    x = torch.randn(2, 32, H, H)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn_like(yr)
    yr.backward(gy)
    with installed() as lib:
        xg = x.clone().requires_grad_(True)
        yg = m(xg)
        yg.backward(gy)
    fused_expected = N == 1 and fam in ("std", "line", "window", "dilation")
    assert ("tramba_ss2d_fused_fwd" in lib.calls) == fused_expected
    assert ("tramba_sscan_fwd" in lib.calls) == (not fused_expected)
    assert "tramba_layernorm2d_fwd" in lib.calls and "tramba_dwconv2d_fwd" in lib.calls
    tol = dict(rtol=1e-4, atol=1e-4 * float(yr.detach().abs().max()))
    np.testing.assert_allclose(yg.detach().numpy(), yr.detach().numpy(), **tol)
    np.testing.assert_allclose(xg.grad.numpy(), xr.grad.numpy(), rtol=1e-4, atol=1e-4 * float(xr.grad.abs().max()))
    for (n, pg), (_, pr) in zip(m.named_parameters(), ref.named_parameters()):
        assert pg.grad is not None, n
        np.testing.assert_allclose(pg.grad.numpy(), pr.grad.numpy(), rtol=1e-3,
                                   atol=1e-4 * float(pr.grad.abs().max()) + 1e-7, err_msg=n)


def test_selective_scan_oflex_contract():
    """11-value backward, fp32 output for bf16 inputs (oflex), None for absent D / delta_bias."""
    g = torch.Generator().manual_seed(1)
    B, KD, L, K, N = 2, 8, 40, 2, 3
    u = torch.randn(B, KD, L, generator=g).bfloat16().requires_grad_(True)
    delta = torch.randn(B, KD, L, generator=g).bfloat16().requires_grad_(True)
    A = (-torch.rand(KD, N, generator=g) - 0.5).requires_grad_(True)
    Bm = torch.randn(B, K, N, L, generator=g).bfloat16().requires_grad_(True)
    Cm = torch.randn(B, K, N, L, generator=g).bfloat16().requires_grad_(True)
    with installed():
        out = tramba_b200.SelectiveScanOflex.apply(u, delta, A, Bm, Cm, None, None, True, -1, -1, True)
        assert out.dtype == torch.float32 and out.shape == (B, KD, L)
        grads = torch.autograd.grad(out, [u, delta, A, Bm, Cm], torch.ones_like(out))
    assert grads[0].dtype == torch.bfloat16 and grads[2].dtype == torch.float32
    ref = OracleSelectiveScan.apply(u.float(), delta.float(), A, Bm.float(), Cm.float(), None, None, True)
    np.testing.assert_allclose(out.detach().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-5)


def test_dct2d_module_and_custom_pair():
    with installed():
        m = tramba_b200.DCT2D(24, 24)
        x = torch.randn(2, 3, 24, 24, requires_grad=True)
        high, low = m(x)
        (high.sum() + 2 * low.sum()).backward()
        tab = so.order_table("hilbert", 12)
        S, M = tramba_b200.make_custom_pair("_h", lambda H: tab)
        z = torch.randn(1, 2, 12, 12)
        assert torch.equal(S.apply(z), tramba_b200.CrossScan_Hilbert.apply(z))
    from oracle import dct as odct
    hi, lo = odct.dct2d_split(x.detach().double().numpy())
    np.testing.assert_allclose(high.detach().numpy(), hi, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(low.detach().numpy(), lo, rtol=1e-5, atol=1e-5)
    assert x.grad is not None and x.grad.shape == x.shape


def test_emulator_scope_is_restored():
    with installed():
        assert torch.zeros(1).is_cuda
    assert not torch.zeros(1).is_cuda
    assert not hasattr(_lib.lib, "calls")


def test_ss2d_module_bf16_marshalling():
    """bf16 activations with fp32 parameters (what autocast hands the operators on the GPU): dtype codes, the fp32
    merged output of the fused core, fp32 dB / dC cast back, LayerNorm2d fp32 -> bf16.  CPU autocast stands in for
    CUDA autocast; the result must stay within bf16 rounding of the fp32 run."""
    torch.manual_seed(2)
    m = tramba_b200.SS2D(d_model=32, d_state=1, channel_first=True, k_group=8, scan=tramba_b200.CrossScan_Line,
                         merge=tramba_b200.CrossMerge_Line)
    x = torch.randn(2, 32, 24, 24)
    with installed() as lib:
        x32 = x.clone().requires_grad_(True)
        y32 = m(x32)
        y32.sum().backward()
        g32 = x32.grad.clone()
        n32 = len(lib.calls)
        x16 = x.clone().requires_grad_(True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            y16 = m(x16)
        y16.float().sum().backward()
    assert "tramba_ss2d_fused_fwd" in lib.calls[n32:] and "tramba_ss2d_fused_bwd" in lib.calls[n32:]
    assert y16.dtype == torch.bfloat16 and x16.grad is not None and torch.isfinite(x16.grad).all()
    rel = (y16.float() - y32).abs().max() / y32.abs().max()
    assert rel < 5e-2, rel
    relg = (x16.grad - g32).abs().max() / g32.abs().max()
    assert relg < 1e-1, relg


SIM_CASES = [("std", "CrossScan", "CrossMerge", 4, 1, 12), ("line", "CrossScan_Line", "CrossMerge_Line", 8, 1, 24),
             ("dilation", "CrossScan_Dilation", "CrossMerge_Dilation", 4, 2, 12),
             ("hilbert", "CrossScan_Hilbert", "CrossMerge_Hilbert", 4, 1, 12)]


@pytest.mark.parametrize("fam,scan,merge,K,N,H", SIM_CASES)
def test_ss2d_module_on_the_simulated_library(fam, scan, merge, K, N, H):
    """Same module test as above, but nothing is emulated by the oracle: the product's Python host calls the product's
    own C ABI (capi.cu) and kernels (fused core or cross-scan / selective scan / cross-merge, proj gather / scatter,
    LayerNorm2d cluster kernels, depthwise conv), compiled as host code and executed on the CPU simulator
    (tools/cpu_sim).  The oracle is only the reference the results are compared with."""
    torch.manual_seed(0)
    m = tramba_b200.SS2D(d_model=32, d_state=N, ssm_ratio=2.0, dt_rank="auto", d_conv=3, conv_bias=False,
                         channel_first=True, disable_z=True, scan=getattr(tramba_b200, scan),
                         merge=getattr(tramba_b200, merge), k_group=K)
    ref = copy.deepcopy(m)
    ref.scan, ref.merge = make_pair(fam)
    ref.selective_scan = OracleSelectiveScan
    ref.fused_dwconv = False
    x = torch.randn(2, 32, H, H)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn_like(yr)
    yr.backward(gy)
    with installed("sim") as lib:
        xg = x.clone().requires_grad_(True)
        yg = m(xg)
        yg.backward(gy)
    fused_expected = N == 1 and fam in ("std", "line")
    assert ("tramba_ss2d_fused_bwd" in lib.calls) == fused_expected
    assert ("tramba_sscan_bwd" in lib.calls) == (not fused_expected)
    for name in ("tramba_layernorm2d_fwd", "tramba_layernorm2d_bwd", "tramba_dwconv2d_fwd", "tramba_dwconv2d_bwd"):
        assert name in lib.calls
    np.testing.assert_allclose(yg.detach().numpy(), yr.detach().numpy(), rtol=1e-3,
                               atol=1e-3 * float(yr.detach().abs().max()))
    np.testing.assert_allclose(xg.grad.numpy(), xr.grad.numpy(), rtol=1e-3, atol=1e-3 * float(xr.grad.abs().max()))
    for (n, pg), (_, pr) in zip(m.named_parameters(), ref.named_parameters()):
        np.testing.assert_allclose(pg.grad.numpy(), pr.grad.numpy(), rtol=2e-3,
                                   atol=2e-3 * float(pr.grad.abs().max()) + 1e-7, err_msg=n)
