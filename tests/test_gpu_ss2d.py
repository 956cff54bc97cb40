"""GPU parity of the composed path (SS2D.forward_core = scan -> proj -> selective scan -> merge,
Models/vmamba.py:231-273) and of FreqSS2D-style use of DCT2D + window/dilated SS2D: the same
module evaluated on CPU with the oracle operators vs on the B200 with the product kernels."""
import copy

import numpy as np
import pytest
import torch

import tramba_b200
from _oracle_ops import OracleSelectiveScan, make_pair

pytestmark = pytest.mark.gpu

CASES = [
    ("std", tramba_b200.CrossScan, tramba_b200.CrossMerge, 4, 1, 24),
    ("line", tramba_b200.CrossScan_Line, tramba_b200.CrossMerge_Line, 8, 1, 24),
    ("window", tramba_b200.CrossScan_Window, tramba_b200.CrossMerge_Window, 4, 1, 24),
    ("dilation", tramba_b200.CrossScan_Dilation, tramba_b200.CrossMerge_Dilation, 4, 1, 12),
    ("std", tramba_b200.CrossScan, tramba_b200.CrossMerge, 4, 16, 12),      # config-2 state size
]


@pytest.mark.parametrize("fam,scan,merge,K,N,H", CASES)
def test_ss2d_module_fwd_bwd_vs_oracle(fam, scan, merge, K, N, H):
    torch.manual_seed(0)
    m = tramba_b200.SS2D(d_model=32, d_state=N, ssm_ratio=2.0, dt_rank="auto", d_conv=3, conv_bias=False,
                         channel_first=True, disable_z=True, scan=scan, merge=merge, k_group=K)
    ref = copy.deepcopy(m)
    ref.scan, ref.merge = make_pair(fam)
    ref.selective_scan = OracleSelectiveScan
    x = torch.randn(2, 32, H, H)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn_like(yr)
    yr.backward(gy)
    m = m.cuda()
    xg = x.cuda().requires_grad_(True)
    yg = m(xg)
    yg.backward(gy.cuda())
    tol = dict(rtol=1e-3, atol=1e-3 * float(yr.abs().max()))
    np.testing.assert_allclose(yg.detach().cpu().numpy(), yr.detach().numpy(), **tol)
    np.testing.assert_allclose(xg.grad.cpu().numpy(), xr.grad.numpy(), rtol=1e-3, atol=1e-3 * float(xr.grad.abs().max()))
    for (n, pg), (_, pr) in zip(m.named_parameters(), ref.named_parameters()):
        assert pg.grad is not None, n
        np.testing.assert_allclose(pg.grad.cpu().numpy(), pr.grad.numpy(), rtol=2e-3,
                                   atol=2e-3 * float(pr.grad.abs().max()) + 1e-7, err_msg=n)


def test_ss2d_bf16_autocast_runs_and_is_close():
    torch.manual_seed(1)
    m = tramba_b200.SS2D(d_model=64, d_state=1, channel_first=True, k_group=8, scan=tramba_b200.CrossScan_Line,
                         merge=tramba_b200.CrossMerge_Line).cuda()
    x = torch.randn(2, 64, 24, 24, device="cuda", requires_grad=True)
    y32 = m(x)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y16 = m(x)
    y16.float().sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
    rel = (y16.float() - y32).abs().max() / y32.abs().max()
    assert rel < 5e-2, rel
