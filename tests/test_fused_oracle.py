"""CPU check of the test-side statement of the fused core (tests/_fused_oracle.py): it must equal
the op-by-op oracle path CrossScan -> grouped projections-free selective scan -> CrossMerge that
is pinned to the reference's golden vectors (tests/test_oracle.py)."""
import numpy as np
import pytest
import torch

from _fused_oracle import fused_core_oracle, param_of
from _oracle_ops import OracleSelectiveScan, make_pair
from oracle import scan_orders as so


@pytest.mark.parametrize("fam,K,H", [("std", 4, 12), ("dilation", 4, 12), ("window", 4, 12), ("line", 8, 12)])
def test_image_order_statement_equals_op_by_op_oracle(fam, K, H):
    g = torch.Generator().manual_seed(H + K)
    B, D, L = 2, 3, H * H
    x = torch.randn(B, D, H, H, generator=g)
    dts = torch.randn(B, K, D, L, generator=g)
    Bs, Cs = torch.randn(B, K, L, generator=g), torch.randn(B, K, L, generator=g)
    A, Ds, bias = -(0.2 + torch.rand(K * D, generator=g)), torch.randn(K * D, generator=g), torch.randn(K * D, generator=g)
    dy = torch.randn(B, D, L, generator=g)
    y, grads = fused_core_oracle(fam, H, [x, dts, Bs, Cs, A, Ds, bias], dy)
    # op-by-op: the pinned operators on the same scan-ordered delta / B / C
    tab = so.order_table(fam, H, param_of(fam, H))
    Scan, Merge = make_pair(fam)
    xr = x.clone().requires_grad_(True)
    xs = Scan.apply(xr)
    Bsc, Csc = Bs.view(B, K, 1, L), Cs.view(B, K, 1, L)
    ys = OracleSelectiveScan.apply(xs.view(B, K * D, L), dts.reshape(B, K * D, L), A.view(-1, 1), Bsc, Csc, Ds, bias,
                                   True, 1, 1, True)
    y2 = Merge.apply(ys.view(B, K, D, H, H))
    y2.backward(dy)
    np.testing.assert_allclose(y.numpy(), y2.detach().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(grads[0].numpy(), xr.grad.numpy(), rtol=1e-4, atol=1e-5)
