"""GPU parity: selective scan fwd/bwd through the C-ABI (via the reference-shaped operator)
vs the float64 oracle.  Tolerances are BASELINE.json's: rtol 1e-3 for fp32 I/O, 1e-2 for bf16;
reductions (dA, dB, dC, dD, ddelta_bias) are compared against the fp64 oracle with an atol
scaled to the magnitude of the summands."""
import numpy as np
import pytest
import torch

import tramba_b200
from tramba_b200 import _lib
from oracle import sscan as osscan

pytestmark = pytest.mark.gpu


def make_inputs(b, K, Dg, L, N, seed=0, dtype=torch.float32, big_delta=False):
    """Synthetic inputs shaped like the reference call site (Models/vmamba.py:244-257) with the
    reference initialisers' ranges (Models/mamba_init.py): A = -(1..N), D = 1, delta_bias =
    softplus^-1(U_log[1e-3, 1e-1])."""
    g = torch.Generator().manual_seed(seed)
    dim = K * Dg
    u = torch.randn(b, dim, L, generator=g)
    delta = torch.randn(b, dim, L, generator=g) * (8.0 if big_delta else 1.0)
    A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(dim, 1) * (0.5 + torch.rand(dim, 1, generator=g))
    B = torch.randn(b, K, N, L, generator=g)
    C = torch.randn(b, K, N, L, generator=g)
    D = 1.0 + 0.1 * torch.randn(dim, generator=g)
    dt = torch.exp(torch.rand(dim, generator=g) * (np.log(0.1) - np.log(1e-3)) + np.log(1e-3))
    bias = dt + torch.log(-torch.expm1(-dt))
    cast = lambda t: t.to(dtype)
    return cast(u), cast(delta), A, cast(B), cast(C), D, bias


def oracle_all(args, dout, softplus=True):
    f = lambda t: None if t is None else t.double().numpy()
    a = [f(t) for t in args]
    out = osscan.selective_scan_fwd(*a, delta_softplus=softplus)
    grads = osscan.selective_scan_bwd(*a, f(dout), delta_softplus=softplus)
    return out, grads


def run_gpu(args, dout, softplus=True, oflex=True):
    dev = [None if t is None else t.cuda().requires_grad_(True) for t in args]
    out = tramba_b200.SelectiveScanOflex.apply(*dev, softplus, -1, -1, oflex)
    grads = torch.autograd.grad(out, [t for t in dev if t is not None], dout.cuda().to(out.dtype))
    it = iter(grads)
    grads = [None if t is None else next(it) for t in dev]
    return out, grads


def check(name, got, ref, rtol, atol_scale=1.0):
    got = got.detach().double().cpu().numpy()
    scale = np.abs(ref).max() + 1e-30
    err = np.abs(got - ref)
    tol = rtol * np.abs(ref) + rtol * atol_scale * scale * 0.05
    bad = err > tol
    assert not bad.any(), f"{name}: {bad.sum()} / {bad.size} off, max err {err.max():.3e} (scale {scale:.3e})"


SHAPES = [
    # (b, K, Dg, L, N)            plan
    (2, 4, 8, 9216, 1),         # TSEG 12 uniform, 24 chunks/row   (stage-0 / freq@96 / dec@96 shape)
    (2, 4, 8, 2304, 1),         # TSEG 12 uniform, 6 chunks/row
    (2, 8, 8, 576, 1),          # TSEG 9 uniform, 2 chunks/row     (helix K = 8)
    (2, 4, 16, 144, 1),         # TSEG 9, two rows per chunk
    (1, 4, 8, 384, 1),          # single chunk rows
    (1, 2, 8, 96, 1),           # TSEG 4, rows straddle chunks
    (1, 2, 8, 196, 1),          # TSEG 4, 14x14
    (1, 2, 3, 49, 1),           # generic scalar path (7x7)
    (1, 2, 5, 50, 2),           # generic scalar path, N = 2
    (2, 4, 8, 2304, 16),        # BASELINE config-2 state size
    (1, 4, 8, 576, 4),
    (1, 1, 8, 36864, 1),        # 192x192 (768^2 stress)
    # rows % R != 0: the last unit is partial and its trailing chunks are empty (generic TMA path)
    (1, 2, 3, 96, 1),
    (1, 1, 5, 96, 2),
    # several rows per chunk with channels-per-group NOT a multiple of rows-per-chunk: sub-rows of one
    # chunk belong to different (batch, group)s -> must not take the side-by-side backward kernel
    (1, 2, 3, 64, 1),
    (1, 2, 3, 64, 2),
    (2, 2, 3, 32, 1),
    (1, 2, 2, 32, 1),
    (1, 2, 4, 64, 1),           # same plan, aligned groups (side-by-side kernel)
]


@pytest.mark.parametrize("b,K,Dg,L,N", SHAPES)
def test_fwd_bwd_fp32(b, K, Dg, L, N):
    args = make_inputs(b, K, Dg, L, N, seed=L + N)
    dout = torch.randn(b, K * Dg, L, generator=torch.Generator().manual_seed(1))
    ref_out, ref_g = oracle_all(args, dout)
    out, g = run_gpu(args, dout)
    assert out.dtype == torch.float32
    check("out", out, ref_out, 1e-3)
    for name, gi, ri in zip("du ddelta dA dB dC dD ddelta_bias".split(), g, ref_g):
        check(name, gi, ri, 1e-3)


@pytest.mark.parametrize("b,K,Dg,L,N", [(2, 4, 8, 2304, 1), (2, 8, 8, 576, 1), (2, 4, 16, 144, 1), (1, 2, 8, 96, 1),
                                         (1, 4, 8, 576, 4)])
def test_fwd_bwd_bf16(b, K, Dg, L, N):
    """bf16 u/delta/B/C (what autocast feeds the operator), fp32 internals and fp32 out (oflex)."""
    args = make_inputs(b, K, Dg, L, N, seed=7, dtype=torch.bfloat16)
    dout = torch.randn(b, K * Dg, L, generator=torch.Generator().manual_seed(2))
    ref_out, ref_g = oracle_all(args, dout)           # oracle sees the same bf16-rounded inputs
    out, g = run_gpu(args, dout)
    assert out.dtype == torch.float32 and g[0].dtype == torch.bfloat16
    check("out", out, ref_out, 1e-3)                  # inputs identical -> fp32-grade output
    for name, gi, ri in zip("du ddelta dA dB dC dD ddelta_bias".split(), g, ref_g):
        check(name, gi, ri, 1e-2)


def test_bf16_out_when_not_oflex():
    args = make_inputs(1, 4, 8, 576, 1, seed=9, dtype=torch.bfloat16)
    dout = torch.randn(1, 32, 576)
    ref_out, _ = oracle_all(args, dout)
    out, g = run_gpu(args, dout.bfloat16(), oflex=False)
    assert out.dtype == torch.bfloat16
    check("out", out, ref_out, 1e-2)


def test_softplus_threshold_and_no_softplus():
    args = make_inputs(1, 2, 8, 576, 1, seed=11, big_delta=True)       # delta+bias well beyond 20 in places
    assert (args[1] + args[6][None, :, None] > 20).any()
    args = list(args)
    args[2] = args[2] * 0.01                                           # keep exp(delta*A) away from underflow
    dout = torch.randn(1, 16, 576)
    ref_out, ref_g = oracle_all(args, dout)
    out, g = run_gpu(args, dout)
    check("out", out, ref_out, 1e-3)
    for name, gi, ri in zip("du ddelta dA dB dC dD ddelta_bias".split(), g, ref_g):
        check(name, gi, ri, 1e-3)
    args2 = list(make_inputs(1, 2, 8, 576, 1, seed=12))
    args2[1] = args2[1].abs() * 0.1
    ref_out, ref_g = oracle_all(args2, dout, softplus=False)
    out, g = run_gpu(args2, dout, softplus=False)
    check("out", out, ref_out, 1e-3)
    for name, gi, ri in zip("du ddelta dA dB dC dD ddelta_bias".split(), g, ref_g):
        check(name, gi, ri, 1e-3)


def test_optional_D_and_bias_none():
    u, delta, A, B, C, D, bias = make_inputs(1, 2, 8, 288, 1, seed=13)
    dout = torch.randn(1, 16, 288)
    for args in ((u, delta, A, B, C, None, bias), (u, delta, A, B, C, D, None), (u, delta, A, B, C, None, None)):
        ref_out, ref_g = oracle_all(args, dout)
        out, g = run_gpu(args, dout)
        check("out", out, ref_out, 1e-3)
        for name, gi, ri in zip("du ddelta dA dB dC dD ddelta_bias".split(), g, ref_g):
            if gi is not None:
                check(name, gi, ri, 1e-3)


def test_inference_without_checkpoints_matches():
    import ctypes
    args = [t.cuda() for t in make_inputs(2, 4, 8, 2304, 1, seed=3)]
    u, delta, A, B, C, D, bias = args
    out1 = tramba_b200.selective_scan_cuda_oflex.fwd(u, delta, A, B, C, D, bias, True, 1, True)[0]
    out2 = torch.empty_like(out1)
    _lib.check(_lib.lib.tramba_sscan_fwd(u.data_ptr(), delta.data_ptr(), A.data_ptr(), B.data_ptr(), C.data_ptr(),
                                         D.data_ptr(), bias.data_ptr(), out2.data_ptr(), None, 2, 32, 2304, 1, 4,
                                         _lib.F32, _lib.F32, 1, None))
    torch.cuda.synchronize()
    assert torch.equal(out1, out2)


def test_full_size_config2_rows_vs_oracle_and_linearity():
    """BASELINE config 2 operator shape (B=4, K*D=1024, L=9216, N=16).  Rows are independent, so a
    random subset of rows is checked against the oracle exactly as a small problem; the whole
    tensor is checked through linearity in u (the recurrence is linear in u for fixed delta)."""
    b, K, Dg, L, N = 4, 4, 256, 9216, 16
    u, delta, A, B, C, D, bias = [t.cuda() for t in make_inputs(b, K, Dg, L, N, seed=5)]
    f = lambda uu: tramba_b200.selective_scan_cuda_oflex.fwd(uu, delta, A, B, C, D, bias, True, 1, True)[0]
    out = f(u)
    u2 = torch.randn_like(u)
    lin = f(u + 2 * u2) - (out + 2 * f(u2))
    assert lin.abs().max() <= 2e-3 * out.abs().max()
    rs = np.random.RandomState(0)
    for _ in range(3):
        bb, k = rs.randint(b), rs.randint(K)
        ds = np.sort(rs.choice(Dg, 4, replace=False)) + k * Dg
        sub = lambda t: t[bb:bb + 1, ds].cpu()
        ref = osscan.selective_scan_fwd(sub(u).double().numpy(), sub(delta).double().numpy(), A[ds].cpu().double().numpy(),
                                        np.repeat(B[bb:bb + 1, k:k + 1].cpu().double().numpy(), 4, 1),
                                        np.repeat(C[bb:bb + 1, k:k + 1].cpu().double().numpy(), 4, 1),
                                        D[ds].cpu().double().numpy(), bias[ds].cpu().double().numpy(), True)
        check("rows", sub(out), ref, 1e-3)


def test_errors_are_loud():
    u, delta, A, B, C, D, bias = [t.cuda() for t in make_inputs(1, 2, 4, 48, 1)]
    with pytest.raises(RuntimeError):
        tramba_b200.selective_scan_cuda_oflex.fwd(u.half(), delta.half(), A, B.half(), C.half(), D, bias, True, 1, True)
    with pytest.raises(RuntimeError):
        tramba_b200.selective_scan_cuda_oflex.fwd(u, delta[:, :4], A, B, C, D, bias, True, 1, True)
    rc = _lib.lib.tramba_sscan_fwd(u.data_ptr(), None, A.data_ptr(), B.data_ptr(), C.data_ptr(), None, None,
                                   u.data_ptr(), None, 1, 8, 48, 1, 2, 0, 0, 1, None)
    assert rc == -1
    rc = _lib.lib.tramba_sscan_fwd(u.data_ptr(), delta.data_ptr(), A.data_ptr(), B.data_ptr(), C.data_ptr(), None,
                                   None, u.data_ptr(), None, 1, 8, 48, 1, 3, 0, 0, 1, None)
    assert rc == -2                                                     # dim % groups


def test_fwd_cross_check_against_vllm_selective_scan():
    """Independent GPU-side cross-check: vLLM's `selective_scan_fn` is a CUDA build of the same
    state-spaces/mamba selective-scan lineage the reference's absent `selective_scan_cuda_oflex` derives from
    (SURVEY 8c).  Forward only (vLLM ships no backward).  Skips cleanly when the op is not importable / has no sm_100
    image.  Same operands (fp32, grouped B/C, D skip, delta_bias, softplus) -> same output to rtol 1e-3."""
    try:
        from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"vLLM selective_scan_fn not importable: {type(e).__name__}")
    b, K, Dg, L, N = 2, 4, 8, 2304, 16
    args = make_inputs(b, K, Dg, L, N, seed=21)
    u, delta, A, B, C, D, bias = [t.cuda() for t in args]
    ours = tramba_b200.SelectiveScanOflex.apply(u, delta, A, B, C, D, bias, True, -1, -1, True)
    try:
        state = torch.zeros(b, K * Dg, N, device="cuda")
        d2 = delta.clone()
        theirs = selective_scan_fn(u.clone(), state, d2, A.contiguous(), B.clone(), C.clone(), D, None, bias, True)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"vLLM selective_scan_fn did not run on this device: {type(e).__name__}: {str(e)[:120]}")
    if theirs is None:
        theirs = d2          # older builds write the output in place of delta
    check("out vs vLLM", ours, theirs.double().cpu().numpy(), 1e-3)
