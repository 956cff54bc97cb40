"""The library's CUDA kernels executed on the CPU (tools/cpu_sim: the .cu sources compiled as host code, every CUDA
thread a cooperative fiber, inline PTX and <<<>>> launches replaced by functional equivalents) against the oracle.

  fused SS2D core fwd / bwd (through the simulated C ABI)  every order family, both dtypes, several items per CTA
  operator selective scan fwd / bwd (sscan_*.cu)          whole files: kernels + host launchers, every plan type
  table-driven cross-scan / cross-merge (cross_scan.cu)   every order family incl. ablation orders and caller tables
  depthwise conv (+ SiLU) fwd / bwd, projected-channel gather / scatter

Covers what can be known without a GPU: index arithmetic, work decomposition over CTAs / rounds / items, the scans
and their carries, barrier and mbarrier protocols (a mismatch is reported as a deadlock), asynchronous-copy
discipline at both extremes (copies landing at issue / only at the wait), and the numerics of the recurrences.  The
kernels are GPU-verified (tests/test_gpu_*.py): here they validate the simulator, and protect later edits.
Interleavings between the two
copy extremes, memory-model ordering and performance are out of its reach.  LayerNorm2d (thread-block clusters,
distributed shared memory) and DCT2D (mma.sync TF32) are covered at the end of this file; the library's whole C ABI on
the simulator is exercised from the Python host in tests/test_host_logic_emulated.py and tests/test_model_parity.py."""
import ctypes
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from tramba_b200 import _lib
from _fused_oracle import fused_core_oracle, param_of

ORDER = {"std": (_lib.ORDER_STD, 4), "dilation": (_lib.ORDER_DILATION, 4), "window": (_lib.ORDER_WINDOW, 4),
         "line": (_lib.ORDER_LINE, 8)}
@pytest.fixture(scope="module")
def sim():
    spec = importlib.util.spec_from_file_location("build_sim", os.path.join(ROOT, "tools", "cpu_sim", "build_sim.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # TRAMBA_SIM_ASAN=1 (with LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0): the
    # AddressSanitizer build -- an out-of-bounds global or shared memory access of a kernel aborts the run
    lib = ctypes.CDLL(mod.build(asan=os.environ.get("TRAMBA_SIM_ASAN", "") == "1"))
    for name in ("tramba_ss2d_fused_fwd", "tramba_ss2d_fused_bwd", "tramba_ss2d_fused_supported"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = _lib.SIGNATURES[name]
    lib.sim_set_sm_count.restype, lib.sim_set_sm_count.argtypes = None, [ctypes.c_int]
    lib.sim_set_lazy.restype, lib.sim_set_lazy.argtypes = None, [ctypes.c_int]
    lib.sim_set_schedule.restype, lib.sim_set_schedule.argtypes = None, [ctypes.c_uint]
    # TRAMBA_SIM_SCHEDULE = 0 (default: ascending thread order) | 1 (descending) | n >= 2 (random permutation per sweep,
    # seed n): every test of this file must pass under every schedule -- a result that depends on the order in which
    # the threads of a barrier interval run is a missing barrier
    lib.sim_set_schedule(int(os.environ.get("TRAMBA_SIM_SCHEDULE", "0")))
    yield lib
    lib.sim_set_lazy(0)
    lib.sim_set_schedule(0)


def _dev(t, bf16):
    """torch tensor -> contiguous numpy buffer as the kernels see it (bf16 as raw uint16)."""
    t = t.contiguous()
    return t.bfloat16().view(torch.int16).numpy().copy() if bf16 else t.float().numpy().copy()


def _host(a, bf16):
    return torch.from_numpy(a.view(np.int16)).view(torch.bfloat16).float() if bf16 else torch.from_numpy(a)


def _run(sim, fam, H, D, B, bf16, grid, seed=0):
    """The fused core through the simulated library's C ABI (tramba_ss2d_fused_fwd / _bwd as include/tramba_b200.h
    declares them) on raw host buffers; `grid` plays the role of the SM count (persistent grids are sized from it)."""
    order, K = ORDER[fam]
    L = H * H
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, D, H, H, generator=g)
    dts = torch.randn(B, K, D, L, generator=g)
    Bs, Cs = torch.randn(B, K, L, generator=g), torch.randn(B, K, L, generator=g)
    A = -(0.2 + torch.rand(K * D, generator=g))
    Ds = 1 + 0.1 * torch.randn(K * D, generator=g)
    bias = torch.randn(K * D, generator=g) - 2.0
    dy = torch.randn(B, D, L, generator=g)
    if bf16:      # round once so the oracle and the kernels start from identical values
        x, dts, Bs, Cs = [t.bfloat16().float() for t in (x, dts, Bs, Cs)]
    param = param_of(fam, H)
    t16 = np.empty((K, L), dtype=np.uint16)
    _lib.check(_lib.lib.tramba_order_table16_skewed(order, H, H, param, t16.ctypes.data))
    n = int(_lib.lib.tramba_order_sched_len(order, H, H, param))
    sched, kinds = np.empty((K, n), np.uint32), np.empty((2 * K,), np.int32)
    _lib.check(_lib.lib.tramba_order_sched(order, H, H, param, sched.ctypes.data, kinds.ctypes.data))
    code = _lib.BF16 if bf16 else _lib.F32
    ins = dict(x=_dev(x, bf16), dts=_dev(dts, bf16), Bs=_dev(Bs, bf16), Cs=_dev(Cs, bf16), A=_dev(A, False),
               Dp=_dev(Ds, False), bias=_dev(bias, False), dy=_dev(dy, False))
    out = dict(y=np.full((B, D, L), np.nan, np.float32), dx=np.full((B, D, L), np.nan, np.float32),
               ddts=np.zeros((B, K, D, L), np.uint16 if bf16 else np.float32),
               dBs=np.full((B, K, L), np.nan, np.float32), dCs=np.full((B, K, L), np.nan, np.float32),
               dA=np.full((K * D,), np.nan, np.float32), dD=np.full((K * D,), np.nan, np.float32),
               dbias=np.full((K * D,), np.nan, np.float32))
    p = lambda a: a.ctypes.data
    sim.sim_set_sm_count(grid)
    rc = sim.tramba_ss2d_fused_fwd(p(ins["x"]), p(ins["dts"]), p(ins["Bs"]), p(ins["Cs"]), p(ins["A"]), p(ins["Dp"]),
                                   p(ins["bias"]), p(t16), p(out["y"]), order, K, B, D, H, H, 1, code, 1, None)
    assert rc == 0, ("fwd", rc)
    have_bwd = bool(sim.tramba_ss2d_fused_supported(order, K, D, H, H, 1, code, 1))
    if have_bwd:
        rc = sim.tramba_ss2d_fused_bwd(p(ins["x"]), p(ins["dts"]), p(ins["Bs"]), p(ins["Cs"]), p(ins["A"]), p(ins["Dp"]),
                                       p(ins["bias"]), p(sched), n, p(kinds), p(ins["dy"]), p(out["dx"]), p(out["ddts"]),
                                       p(out["dBs"]), p(out["dCs"]), p(out["dA"]), p(out["dD"]), p(out["dbias"]), order,
                                       K, B, D, H, H, 1, code, 1, None)
        assert rc == 0, ("bwd", rc)
    ref_y, ref_g = fused_core_oracle(fam, H, [x, dts, Bs, Cs, A, Ds, bias], dy)
    return out, have_bwd, ref_y, ref_g


def _close(a, b, rtol, what):
    a, b = np.asarray(a, np.float64), b.detach().double().numpy()
    tol = rtol * np.abs(b) + 0.05 * rtol * np.abs(b).max()
    bad = np.abs(a - b) > tol
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} off, max err {np.abs(a - b).max():.3e}"


CASES = [  # family, H, D, B, bf16, grid
    ("std", 12, 64, 2, False, 3), ("std", 12, 128, 1, True, 2), ("line", 12, 64, 1, False, 3),
    ("window", 24, 32, 2, False, 3), ("dilation", 24, 16, 1, True, 1), ("line", 24, 16, 1, True, 2),
    ("std", 48, 8, 1, False, 2), ("window", 48, 4, 1, True, 1), ("std", 96, 2, 1, True, 2),
    ("dilation", 96, 8, 1, True, 3),   # 96 x 96, one plane per round, several rounds per backward item and items per CTA
    ("line", 96, 2, 1, True, 2),       # helix duplicates (shared-memory atomics) with 512 threads per plane
    ("std", 96, 2, 1, False, 2),       # fp32 at 96 x 96: forward only (the backward's staging does not fit)
]


@pytest.mark.parametrize("lazy", [0, 1], ids=["eager-copies", "lazy-copies"])
@pytest.mark.parametrize("fam,H,D,B,bf16,grid", CASES)
def test_fused_kernels_on_the_cpu_simulator(sim, fam, H, D, B, bf16, grid, lazy):
    """lazy = 0: asynchronous copies land when issued (a buffer refilled too early corrupts its readers);
    lazy = 1: they land when waited for and their destinations are poisoned meanwhile, bulk stores / reductions read
    their shared source at the wait_group (a read before the barrier / a staging buffer overwritten too early
    corrupts data)."""
    sim.sim_set_lazy(lazy)
    out, have_bwd, ref_y, ref_g = _run(sim, fam, H, D, B, bf16, grid)
    rtol = 1e-2 if bf16 else 1e-3
    _close(out["y"], ref_y, 1e-3, "y")
    assert have_bwd            # every Tramba-V geometry, both dtypes (fp32 at 96 x 96 reads B / C from global memory)
    if have_bwd:
        _close(out["dx"].reshape(ref_g[0].shape), ref_g[0], rtol, "dx")
        _close(_host(out["ddts"], bf16).numpy(), ref_g[1], rtol, "ddelta")
        _close(out["dBs"], ref_g[2], rtol, "dB")
        _close(out["dCs"], ref_g[3], rtol, "dC")
        _close(out["dA"], ref_g[4], rtol, "dA")
        _close(out["dD"], ref_g[5], rtol, "dD")
        _close(out["dbias"], ref_g[6], rtol, "dbias")


def test_simulator_selftest(sim):
    """The runtime itself: shuffles / segmented scans give the right values, a thread that skips a barrier and an
    mbarrier whose transaction never completes are both REPORTED (as a deadlock) instead of hanging or passing."""
    cp, fp = ctypes.c_char_p, ctypes.c_void_p
    sim.sim_selftest_shuffle.restype, sim.sim_selftest_shuffle.argtypes = cp, [fp, fp, ctypes.c_int]
    sim.sim_selftest_mismatch.restype, sim.sim_selftest_mismatch.argtypes = cp, []
    sim.sim_selftest_mbar.restype, sim.sim_selftest_mbar.argtypes = cp, [ctypes.c_int, fp]
    sim.sim_set_lazy(0)
    x = np.arange(64, dtype=np.float32) + 1
    out = np.zeros(2 * 64 + 1, np.float32)
    assert sim.sim_selftest_shuffle(x.ctypes.data, out.ctypes.data, 3) is None
    assert np.array_equal(out[:32], np.full(32, x[:32].sum())) and np.array_equal(out[32:64], np.full(32, x[32:].sum()))
    want = np.concatenate([np.cumsum(x[i:i + 16]) for i in range(0, 64, 16)])
    assert np.array_equal(out[64:128], want) and out[128] == x[32:].sum() + 2       # thread 63 holds warp 1's sum; the last CTA is blockIdx 2
    assert b"deadlock" in sim.sim_selftest_mismatch()
    o = np.zeros(32, np.float32)
    assert sim.sim_selftest_mbar(1, o.ctypes.data) is None and np.array_equal(o[:4], [1, 2, 3, 4])
    assert b"deadlock" in sim.sim_selftest_mbar(0, o.ctypes.data)


# ------------------------------------------------------------------------------------------------------------------
# operator selective scan (csrc/sscan_fwd.cu, sscan_bwd.cu): the whole files are translated -- kernels AND host
# launchers (plan selection, template dispatch), only the <<<>>> launches / PTX / shared-memory declaration rewritten
SSCAN_CASES = [  # batch, dim, L, N, groups, bf16     (plans: see make_sscan_plan in csrc/sscan.cuh)
    (2, 16, 576, 1, 4, False),     # the model's shape class: N = 1, TMA ring, 2 chunks of 288 per row
    (1, 8, 2304, 1, 2, True),      # bf16 in, fp32 out, 384-token chunks
    (2, 16, 144, 1, 4, False),     # two 144-token rows per chunk (regular, rpc = 2)
    (1, 8, 1152, 16, 4, False),    # config-2 state size N = 16
    (1, 6, 50, 3, 3, False),       # rows that are not 16-byte multiples: the DIRECT scalar path, N = 3
    (2, 8, 9216, 1, 4, True),      # 96 x 96 rows
]


@pytest.mark.parametrize("lazy", [0, 1], ids=["eager-copies", "lazy-copies"])
@pytest.mark.parametrize("batch,dim,L,N,groups,bf16", SSCAN_CASES)
def test_operator_selective_scan_on_the_cpu_simulator(sim, batch, dim, L, N, groups, bf16, lazy):
    from oracle import sscan as osscan
    vp, fp, ci = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int
    sim.sim_sscan_ckpt_elems.restype, sim.sim_sscan_ckpt_elems.argtypes = ctypes.c_size_t, [ci] * 5
    sim.sim_sscan_fwd.restype, sim.sim_sscan_fwd.argtypes = ctypes.c_char_p, [vp] * 9 + [ci] * 9
    sim.sim_sscan_bwd.restype, sim.sim_sscan_bwd.argtypes = ctypes.c_char_p, [vp] * 16 + [ci] * 9
    sim.sim_set_lazy(lazy)
    g = torch.Generator().manual_seed(L + N)
    u, delta = torch.randn(batch, dim, L, generator=g), torch.randn(batch, dim, L, generator=g)
    A = -(0.3 + torch.rand(dim, N, generator=g))
    Bm, Cm = torch.randn(batch, groups, N, L, generator=g), torch.randn(batch, groups, N, L, generator=g)
    D, bias = 1 + 0.1 * torch.randn(dim, generator=g), torch.randn(dim, generator=g) - 2.0
    dout = torch.randn(batch, dim, L, generator=g)
    if bf16:
        u, delta, Bm, Cm = [t.bfloat16().float() for t in (u, delta, Bm, Cm)]
    code = _lib.BF16 if bf16 else _lib.F32
    hu, hd, hB, hC = _dev(u, bf16), _dev(delta, bf16), _dev(Bm, bf16), _dev(Cm, bf16)
    hA, hD, hb, hdo = _dev(A, False), _dev(D, False), _dev(bias, False), _dev(dout, False)
    out = np.full((batch, dim, L), np.nan, np.float32)
    ckpt = np.zeros(max(int(sim.sim_sscan_ckpt_elems(batch, dim, L, N, code)), 1), np.float32)
    p = lambda a: a.ctypes.data
    err = sim.sim_sscan_fwd(p(hu), p(hd), p(hA), p(hB), p(hC), p(hD), p(hb), p(out), p(ckpt), batch, dim, L, N, groups,
                            code, _lib.F32, 1, 3)
    assert err is None, err.decode()
    f = lambda t: t.double().numpy()
    ref = osscan.selective_scan_fwd(f(u), f(delta), f(A), f(Bm), f(Cm), f(D), f(bias), True)
    rtol = 1e-2 if bf16 else 1e-3
    _close(out, torch.from_numpy(ref), rtol, "out")
    du = np.zeros((batch, dim, L), np.uint16 if bf16 else np.float32)
    dd = np.zeros_like(du)
    dA, dB, dC = np.zeros((dim, N), np.float32), np.zeros((batch, groups, N, L), np.float32), \
        np.zeros((batch, groups, N, L), np.float32)
    dD, db = np.zeros(dim, np.float32), np.zeros(dim, np.float32)
    err = sim.sim_sscan_bwd(p(hu), p(hd), p(hA), p(hB), p(hC), p(hD), p(hb), p(hdo), p(ckpt), p(du), p(dd), p(dA), p(dB),
                            p(dC), p(dD), p(db), batch, dim, L, N, groups, code, _lib.F32, 1, 3)
    assert err is None, err.decode()
    refg = osscan.selective_scan_bwd(f(u), f(delta), f(A), f(Bm), f(Cm), f(D), f(bias), f(dout), True)
    got = [_host(du, bf16).numpy(), _host(dd, bf16).numpy(), dA, dB, dC, dD, db]
    for name, a, b in zip(("du", "ddelta", "dA", "dB", "dC", "dD", "dbias"), got, refg):
        _close(a, torch.from_numpy(np.asarray(b)), 2 * rtol, name)


# ------------------------------------------------------------------------------------------------------------------
# table-driven cross-scan / cross-merge (csrc/cross_scan.cu, translated whole): every order family, including the
# ablation orders whose merge configuration (permutation vs CSR path) was added without a GPU at hand
SCAN_FAMS = {"std": (_lib.ORDER_STD, 0), "dilation": (_lib.ORDER_DILATION, 4), "window": (_lib.ORDER_WINDOW, None),
             "line": (_lib.ORDER_LINE, 0), "line4": (_lib.ORDER_LINE4, 0), "spiral": (_lib.ORDER_SPIRAL, 0),
             "hilbert": (_lib.ORDER_HILBERT, 0), "diagonal": (_lib.ORDER_DIAGONAL, 0), "ds": (_lib.ORDER_DS, 0),
             "ab2": (_lib.ORDER_AB2, 0), "ab1": (_lib.ORDER_AB1, 0), "custom": (_lib.ORDER_CUSTOM, 0)}


SCAN_CASES = [(f, H, B, C, bf) for f in SCAN_FAMS for (H, B, C, bf) in ((12, 2, 3, False), (24, 1, 5, True),
                                                                         (96, 1, 2, False))]
SCAN_CASES += [(f, 192, 1, 1, False) for f in ("dilation", "line", "ds")]   # planes too large to stage: the direct kernel


@pytest.mark.parametrize("fam,H,B,C,bf16", SCAN_CASES)
def test_cross_scan_merge_on_the_cpu_simulator(sim, fam, H, B, C, bf16):
    """Bit-exact against the oracle (gather; sequential fp32 sums in the reference's association order for
    std / dilation / window / line, sequential over k for the ablation and custom orders)."""
    from oracle import scan_orders as so
    from tramba_b200 import ops
    vp, ci = ctypes.c_void_p, ctypes.c_int
    for fn in (sim.sim_cross_scan, sim.sim_cross_merge):
        fn.restype, fn.argtypes = ctypes.c_char_p, [vp, vp, vp] + [ci] * 7
    sim.sim_set_lazy(0)
    order, param = SCAN_FAMS[fam]
    if fam == "window":
        param = {12: 4, 24: 8, 96: 16}[H]
    L = H * H
    rs = np.random.RandomState(H + C)
    if H == 192:
        ops.line_sizes.append(192) if 192 not in ops.line_sizes else None
    if fam == "custom":
        tab = rs.randint(0, L, size=(3, L)).astype(np.int64)            # arbitrary table with repeats
        K, plan_t = ops._custom_plan(tab, "cpu")
        plan = plan_t.numpy()
    else:
        tab = so.order_table(fam, H, param)
        K, plan = ops.order_plan_host(order, H, param)
    x = torch.from_numpy(rs.standard_normal((B, C, H, H)).astype(np.float32))
    ys = torch.from_numpy(rs.standard_normal((B, K, C, L)).astype(np.float32))
    if bf16:
        x, ys = x.bfloat16().float(), ys.bfloat16().float()
    code = _lib.BF16 if bf16 else _lib.F32
    hx, hys = _dev(x, bf16), _dev(ys, bf16)
    xs = np.zeros((B, K, C, L), hx.dtype)
    y = np.zeros((B, C, L), hx.dtype)
    err = sim.sim_cross_scan(hx.ctypes.data, xs.ctypes.data, plan.ctypes.data, order, K, B, C, H, H, code)
    assert err is None, err.decode()
    assert np.array_equal(_host(xs, bf16).numpy(), so.cross_scan(x.numpy(), tab))
    err = sim.sim_cross_merge(hys.ctypes.data, y.ctypes.data, plan.ctypes.data, order, K, B, C, H, H, code)
    assert err is None, err.decode()
    want = so.cross_merge(ys.numpy(), tab, fam if fam in ("std", "dilation", "window", "line") else "custom")
    if bf16:
        want = torch.from_numpy(want).bfloat16().float().numpy()           # fp32 sums, one rounding
    assert np.array_equal(_host(y, bf16).numpy(), want)


# ------------------------------------------------------------------------------------------------------------------
# either side of the core: depthwise conv (+ SiLU) and the projected-channel gather / scatter
@pytest.mark.parametrize("B,C,H,k,act,bf16", [(2, 6, 24, 3, "silu", False), (1, 5, 12, 3, "silu", True),
                                              (1, 4, 48, 7, "none", False), (2, 3, 96, 3, "silu", False),
                                              (1, 2, 24, 5, "silu", False)])
def test_dwconv_on_the_cpu_simulator(sim, B, C, H, k, act, bf16):
    from oracle import dwconv as odw
    vp, ci = ctypes.c_void_p, ctypes.c_int
    sim.sim_dwconv2d_fwd.restype, sim.sim_dwconv2d_fwd.argtypes = ctypes.c_char_p, [vp] * 4 + [ci] * 7
    sim.sim_dwconv2d_bwd.restype, sim.sim_dwconv2d_bwd.argtypes = ctypes.c_char_p, [vp] * 6 + [ci] * 7
    sim.sim_set_lazy(0)
    g = torch.Generator().manual_seed(H + k)
    x, dy = torch.randn(B, C, H, H, generator=g), torch.randn(B, C, H, H, generator=g)
    w, b = 0.3 * torch.randn(C, 1, k, k, generator=g), 0.1 * torch.randn(C, generator=g)
    if bf16:
        x, dy = x.bfloat16().float(), dy.bfloat16().float()
    code, actc = (_lib.BF16 if bf16 else _lib.F32), (_lib.ACT_SILU if act == "silu" else _lib.ACT_NONE)
    hx, hdy, hw, hb = _dev(x, bf16), _dev(dy, bf16), _dev(w, False), _dev(b, False)
    y, dx = np.zeros_like(hx), np.zeros_like(hx)
    part = np.full((B, C, k * k + 1), np.nan, np.float32)
    p = lambda a: a.ctypes.data
    err = sim.sim_dwconv2d_fwd(p(hx), p(hw), p(hb), p(y), B, C, H, H, k, actc, code)
    assert err is None, err.decode()
    err = sim.sim_dwconv2d_bwd(p(hx), p(hw), p(hb), p(hdy), p(dx), p(part), B, C, H, H, k, actc, code)
    assert err is None, err.decode()
    f = lambda t: t.double().numpy()
    rtol = 1e-2 if bf16 else 1e-4
    _close(_host(y, bf16).numpy(), torch.from_numpy(odw.dwconv_fwd(f(x), f(w), f(b), act)), rtol, "y")
    gx, gw, gb = odw.dwconv_bwd(f(dy), f(x), f(w), f(b), act)
    _close(_host(dx, bf16).numpy(), torch.from_numpy(gx), rtol, "dx")
    _close(part.sum(0)[:, :k * k].reshape(C, 1, k, k), torch.from_numpy(gw), 10 * rtol, "dw")
    _close(part.sum(0)[:, k * k], torch.from_numpy(gb), 10 * rtol, "db")


@pytest.mark.parametrize("fam,H,R,N,bf16", [("std", 12, 2, 1, False), ("line", 24, 2, 1, True),
                                            ("window", 24, 4, 2, False), ("dilation", 96, 8, 1, False)])
def test_proj_gather_scatter_on_the_cpu_simulator(sim, fam, H, R, N, bf16):
    from oracle import scan_orders as so
    from tramba_b200 import ops
    vp, ci = ctypes.c_void_p, ctypes.c_int
    sim.sim_proj_gather.restype, sim.sim_proj_gather.argtypes = ctypes.c_char_p, [vp] * 5 + [ci] * 6
    sim.sim_proj_scatter.restype, sim.sim_proj_scatter.argtypes = ctypes.c_char_p, [vp] * 5 + [ci] * 7
    sim.sim_set_lazy(0)
    order, param = SCAN_FAMS[fam]
    if fam == "window":
        param = 8
    L, B = H * H, 2
    tab = so.order_table(fam, H, param)
    K, plan = ops.order_plan_host(order, H, param)
    CX = R + 2 * N
    rs = np.random.RandomState(H)
    src = torch.from_numpy(rs.randint(-8, 9, size=(B, K, CX, L)).astype(np.float32))      # exact in bf16 too
    hsrc = _dev(src, bf16)
    code = _lib.BF16 if bf16 else _lib.F32
    dts, Bs, Cs = (np.zeros((B, K, n, L), hsrc.dtype) for n in (R, N, N))
    p = lambda a: a.ctypes.data
    err = sim.sim_proj_gather(p(hsrc), p(dts), p(Bs), p(Cs), p(plan), K, B, R, N, L, code)
    assert err is None, err.decode()
    want = np.stack([src.numpy()[:, k][:, :, tab[k]] for k in range(K)], axis=1)
    got = np.concatenate([_host(a, bf16).numpy() for a in (dts, Bs, Cs)], axis=2)
    assert np.array_equal(got, want)
    # adjoint: scatter(gather-shaped gradients) == transpose, exact on small integers; fp32 B / C gradients
    gB, gC = (rs.randint(-8, 9, size=(B, K, N, L)).astype(np.float32) for _ in range(2))
    gsrc = np.zeros_like(hsrc)
    err = sim.sim_proj_scatter(p(dts), p(gB), p(gC), p(gsrc), p(plan), K, B, R, N, L, code, _lib.F32)
    assert err is None, err.decode()
    g = np.concatenate([_host(dts, bf16).numpy(), gB, gC], axis=2)
    ref = np.zeros((B, K, CX, L), np.float32)
    for k in range(K):
        np.add.at(ref[:, k], (slice(None), slice(None), tab[k]), g[:, k])
    if bf16:
        ref = torch.from_numpy(ref).bfloat16().float().numpy()
    assert np.array_equal(_host(gsrc, bf16).numpy(), ref)


# ------------------------------------------------------------------------------------------------------------------
# LayerNorm2d (+ GELU): the register-resident CLUSTER kernels (cluster.sync, distributed shared memory) and the
# streaming fallback; DCT2D on the tensor pipe (mma.sync m16n8k8 TF32 emulated fragment by fragment) and SIMT
@pytest.mark.parametrize("B,C,L,act,in_bf16,out_bf16,sms", [
    (2, 32, 576, "gelu", False, False, 3),      # few channels: one CTA per tile
    (1, 256, 144, "none", False, True, 148),    # many channels, small map: cluster of CTAs over the channels (DSMEM)
    (2, 128, 2304, "gelu", True, False, 8),
    (1, 64, 50, "none", False, False, 3),       # odd L: streaming kernels
    (1, 512, 144, "gelu", False, False, 148),
])
def test_layernorm2d_on_the_cpu_simulator(sim, B, C, L, act, in_bf16, out_bf16, sms):
    from oracle import layernorm2d as oln
    vp, ci = ctypes.c_void_p, ctypes.c_int
    sim.sim_layernorm2d_tiles.restype, sim.sim_layernorm2d_tiles.argtypes = ci, [ci] * 4
    sim.sim_layernorm2d_fwd.restype = ctypes.c_char_p
    sim.sim_layernorm2d_fwd.argtypes = [vp] * 6 + [ci] * 3 + [ctypes.c_float] + [ci] * 4
    sim.sim_layernorm2d_bwd.restype, sim.sim_layernorm2d_bwd.argtypes = ctypes.c_char_p, [vp] * 9 + [ci] * 7
    sim.sim_set_lazy(0)
    g = torch.Generator().manual_seed(C + L)
    x = torch.randn(B, C, L, generator=g) * 2 + 0.5
    gy = torch.randn(B, C, L, generator=g)
    w, b = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    if in_bf16:
        x = x.bfloat16().float()
    actc = _lib.ACT_GELU if act == "gelu" else _lib.ACT_NONE
    ic, oc = (_lib.BF16 if in_bf16 else _lib.F32), (_lib.BF16 if out_bf16 else _lib.F32)
    hx, hw, hb, hgy = _dev(x, in_bf16), _dev(w, False), _dev(b, False), _dev(gy, False)
    y = np.zeros((B, C, L), np.uint16 if out_bf16 else np.float32)
    mean, rstd = np.zeros((B, L), np.float32), np.zeros((B, L), np.float32)
    p = lambda a: a.ctypes.data
    err = sim.sim_layernorm2d_fwd(p(hx), p(hw), p(hb), p(y), p(mean), p(rstd), B, C, L, 1e-5, actc, ic, oc, sms)
    assert err is None, err.decode()
    f = lambda t: t.double().numpy()
    _close(_host(y, out_bf16).numpy(), torch.from_numpy(oln.layernorm2d_fwd(f(x), f(w), f(b), 1e-5, act)),
           1e-2 if out_bf16 else 1e-4, "y")
    tiles = sim.sim_layernorm2d_tiles(B, C, L, sms)
    assert tiles > 0
    gx = np.zeros_like(hx)
    gw, gb = np.full((B * tiles, C), np.nan, np.float32), np.full((B * tiles, C), np.nan, np.float32)
    err = sim.sim_layernorm2d_bwd(p(hgy), p(hx), p(hw), p(hb), p(mean), p(rstd), p(gx), p(gw), p(gb), B, C, L, actc, ic,
                                  _lib.F32, sms)
    assert err is None, err.decode()
    rx, rw, rb = oln.layernorm2d_bwd(f(gy), f(x), f(w), f(b), 1e-5, act)
    _close(_host(gx, in_bf16).numpy(), torch.from_numpy(rx), 1e-2 if in_bf16 else 2e-4, "gx")
    _close(gw.sum(0), torch.from_numpy(rw), 1e-3, "gw")
    _close(gb.sum(0), torch.from_numpy(rb), 1e-3, "gb")
    sim.sim_last_cluster_size.restype = ci
    cs = sim.sim_last_cluster_size()
    if (C, L, sms) in ((256, 144, 148), (512, 144, 148)):
        assert cs > 1, "expected the cluster / DSMEM kernels for many channels on a small map"
    if L % 2:
        assert cs == 1            # odd L: streaming kernels


@pytest.mark.parametrize("planes,H,tc", [(3, 24, 1), (2, 48, 1), (1, 96, 1), (2, 12, 0), (3, 24, 0)])
def test_dct2d_on_the_cpu_simulator(sim, planes, H, tc):
    from oracle import dct as odct
    vp, ci = ctypes.c_void_p, ctypes.c_int
    sim.sim_dct2d.restype, sim.sim_dct2d.argtypes = ctypes.c_char_p, [ci, ci] + [vp] * 6 + [ci] * 3
    for lazy in (0, 1):          # the tensor-pipe kernel double-buffers its planes with cp.async groups
        sim.sim_set_lazy(lazy)
        rs = np.random.RandomState(H)
        x = rs.standard_normal((1, planes, H, H)).astype(np.float32)
        w = np.empty((H, H), np.float32)
        _lib.lib.tramba_dct_basis(H, w.ctypes.data)
        hi, lo = np.zeros((planes, H // 2, H // 2), np.float32), np.zeros((planes, H // 2, H // 2), np.float32)
        p = lambda a: a.ctypes.data
        err = sim.sim_dct2d(0, tc, p(x), None, p(w), p(w), p(hi), p(lo), planes, H, H)
        assert err is None, err.decode()
        rhi, rlo = odct.dct2d_split(x.astype(np.float64))
        _close(hi, torch.from_numpy(np.ascontiguousarray(rhi[0])), 1e-5, "high")
        _close(lo, torch.from_numpy(np.ascontiguousarray(rlo[0])), 1e-5, "low")
        gh, gl = rs.standard_normal(hi.shape).astype(np.float32), rs.standard_normal(hi.shape).astype(np.float32)
        gx = np.zeros((planes, H, H), np.float32)
        err = sim.sim_dct2d(1, tc, p(gh), p(gl), p(w), p(w), p(gx), None, planes, H, H)
        assert err is None, err.decode()
        ref = odct.dct2d_split_bwd(gh[None].astype(np.float64), gl[None].astype(np.float64), H, H)
        _close(gx, torch.from_numpy(np.ascontiguousarray(ref[0])), 1e-5, "gx")


def test_schedule_shuffling_exposes_a_missing_barrier(sim):
    """The detector behind TRAMBA_SIM_SCHEDULE: a kernel that reads a neighbour's shared-memory slot without a barrier
    gives schedule-dependent results; the library's kernels must not (the whole file is run under schedules 1, 7, 12345
    as well as the default -- see profiles/r01_offline_gpu_tests_on_simulator.txt)."""
    sim.sim_selftest_race.restype, sim.sim_selftest_race.argtypes = ctypes.c_char_p, [ctypes.c_void_p]
    outs = []
    for mode in (0, 1, 7):
        sim.sim_set_schedule(mode)
        o = np.zeros(64, np.float32)
        assert sim.sim_selftest_race(o.ctypes.data) is None
        outs.append(o.copy())
    sim.sim_set_schedule(int(os.environ.get("TRAMBA_SIM_SCHEDULE", "0")))
    assert not (np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2]))
