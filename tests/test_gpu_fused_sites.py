"""GPU parity of the fused SS2D core AT THE BENCHMARKED GEOMETRY: the ten Tramba-V call sites of SURVEY 3.3
(Models/vmamba.py:231-259 as called from VSSBlock / FreqSS2Dv6 / MultiScaleDecoderBlock) with the batch of the
training step (B = 4) and the FULL channel count, so every persistent CTA walks several forward rounds and several
backward items (the buffer-reuse / mbarrier-phase / async-store paths that the small cases never reach).

The fp64 oracle cannot afford the full tensors, and does not need to: planes are independent.
  run 1 (dense dy): y, dx, ddelta and the per-row sums dA / dD / ddelta_bias are compared on a random SUBSET of the
        channel planes (all batch entries, all directions), the oracle evaluated on that subset only;
  run 2 (dy zero outside the subset): dB / dC -- sums over ALL channels of a (batch, direction) -- then receive
        contributions from the subset only (linearity), so the same subset oracle pins them while every CTA still
        executes the full-size schedule.
Tolerances are BASELINE.json's: rtol 1e-3 (fp32) / 1e-2 (bf16 outputs), absolute floor 0.05 * rtol * max|ref|."""
import os

import numpy as np
import pytest
import torch

from tramba_b200 import _lib, fused
from _fused_oracle import fused_core_oracle as _oracle, param_of as _param

pytestmark = pytest.mark.gpu

FAM = {"std": (_lib.ORDER_STD, 4), "dilation": (_lib.ORDER_DILATION, 4), "window": (_lib.ORDER_WINDOW, 4),
       "line": (_lib.ORDER_LINE, 8)}
# (family, H, d_inner) of SURVEY 3.3: encoder stages, the two FreqSS2Dv6 branches, the helix decoder blocks
SITES = [("std", 96, 256), ("std", 48, 512), ("std", 24, 1024), ("std", 12, 2048),
         ("window", 96, 256), ("window", 48, 512), ("window", 24, 1024),
         ("dilation", 96, 256), ("dilation", 48, 512), ("dilation", 24, 1024),
         ("line", 96, 256), ("line", 48, 512), ("line", 24, 1024)]
# on the CPU simulator (tests/conftest.py) the same test logic runs on a 16x thinner problem
_SIM = os.environ.get("TRAMBA_GPU_TESTS_ON_SIM", "") == "1"
BATCH = 2 if _SIM else 4
NSUB = 10
DX_FLOOR = 0.05


def _close(a, b, rtol, what, floor=0.05):
    a = a.detach().double().cpu().numpy()
    b = b.detach().double().cpu().numpy()
    tol = rtol * np.abs(b) + floor * rtol * np.abs(b).max()
    err = np.abs(a - b)
    assert (err <= tol).all(), (f"{what}: {(err > tol).sum()} / {err.size} off, max err {err.max():.3e} "
                                f"(ref max {np.abs(b).max():.3e}) at {np.unravel_index(np.argmax(err - tol), err.shape)}")


def _site_inputs(fam, H, D, dtype, seed):
    order, K = FAM[fam]
    L = H * H
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    x, dts = r(BATCH, D, H, H), r(BATCH, K, D, L)
    Bs, Cs = r(BATCH, K, L), r(BATCH, K, L)
    A = -(0.2 + torch.rand(K * D, generator=g).cuda())
    Ds = 1 + 0.1 * r(K * D)
    bias = r(K * D) - 2.0
    if dtype != torch.float32:   # round once: the oracle sees exactly what the kernel sees
        x, dts, Bs, Cs = (t.to(dtype).float() for t in (x, dts, Bs, Cs))
    return order, K, [x, dts, Bs, Cs, A, Ds, bias]


def _subset(D, seed):
    rs = np.random.RandomState(seed)
    pick = set(rs.choice(D, size=NSUB - 3, replace=False).tolist()) | {0, D // 2 + 1, D - 1}
    return torch.tensor(sorted(pick))


def _take(tensors, K, D, sub):
    x, dts, Bs, Cs, A, Ds, bias = [t.detach().cpu() for t in tensors]
    rows = lambda v: v.view(K, D)[:, sub].reshape(-1)
    return [x[:, sub], dts[:, :, sub], Bs, Cs, rows(A), rows(Ds), rows(bias)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
@pytest.mark.parametrize("fam,H,D", SITES)
def test_fused_core_at_training_geometry(fam, H, D, dtype):
    if _SIM:
        D = max(D // 16, 16)
    order, K, tensors = _site_inputs(fam, H, D, dtype, seed=3 * H + D)
    L = H * H
    assert fused.supported(order, K, D, H, H, 1, dtype, need_bwd=False)
    have_bwd = fused.supported(order, K, D, H, H, 1, dtype, need_bwd=True)
    sub = _subset(D, seed=H)
    dev = [t.to(dtype) if i < 4 else t for i, t in enumerate(tensors)]
    dev = [t.requires_grad_(True) for t in dev]
    dy = torch.randn(BATCH, D, L, generator=torch.Generator().manual_seed(11)).cuda()
    rt_out = 1e-3
    rt_g = 1e-3 if dtype == torch.float32 else 1e-2
    # dx is accumulated over the K directions in fp32 inside the backward kernel and rounded to bf16 ONCE, so it
    # meets the same floor as everything else (the first-generation path summed K bf16-rounded du terms)
    fl_dx = DX_FLOOR

    y = fused.ss2d_core_fused(*dev, order, _param(fam, H))
    yr, gr = _oracle(fam, H, _take(tensors, K, D, sub), dy[:, sub].cpu())
    _close(y[:, sub], yr, rt_out, "y")
    assert have_bwd                  # every Tramba-V site, both dtypes (fp32 at 96 x 96 reads B / C from global memory)
    # ---- run 1: dense dy -> plane-local quantities on the subset
    grads = torch.autograd.grad(y, dev, dy, retain_graph=True)
    dx, ddts, _, _, dA, dD, db = grads
    rows = lambda v: v.view(K, D)[:, sub].reshape(-1)
    _close(dx[:, sub], gr[0], rt_g, "dx", fl_dx)
    _close(ddts[:, :, sub], gr[1], rt_g, "ddts")
    _close(rows(dA), gr[4], rt_g, "dA")
    _close(rows(dD), gr[5], rt_g, "dD")
    _close(rows(db), gr[6], rt_g, "dbias")
    # ---- run 2: dy restricted to the subset -> the cross-channel sums dB / dC
    mask = torch.zeros(D)
    mask[sub] = 1
    mask = mask.cuda()
    grads2 = torch.autograd.grad(y, dev, dy * mask.view(1, D, 1))
    _close(grads2[2], gr[2], rt_g, "dBs")
    _close(grads2[3], gr[3], rt_g, "dCs")
    _close(grads2[0][:, sub], gr[0], rt_g, "dx (masked run)", fl_dx)
    off = torch.ones(D, dtype=torch.bool)
    off[sub] = False
    assert float(grads2[1].cpu()[:, :, off].abs().max()) == 0.0   # zero dy -> exactly zero ddelta
