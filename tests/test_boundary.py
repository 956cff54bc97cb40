"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly what
include/tramba_b200.h declares, its host-side scan-order generators are bit-exact with the
reference's golden tables, and the product never routes through the oracle or a CPU path."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

import tramba_b200
from tramba_b200 import _lib, ops

WS = {12: 4, 24: 8, 48: 12, 96: 16, 192: 32}


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "tramba_b200.h")).read()
    return sorted(set(re.findall(r"TRAMBA_API[^;(]*?\b(tramba_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 16
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == syms                      # nothing missing, nothing extra
    assert sorted(_lib.SIGNATURES) == syms       # the ctypes binding covers the whole header
    assert _lib.lib.tramba_abi_version() == 1


def test_library_is_sm100a_native_code():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", out))
    assert archs == {"sm_100a"}, archs


def test_tma_bulk_copies_present_in_sass():
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in out            # cp.async.bulk (TMA 1-D) in SASS
    assert "SYNCS" in out             # mbarrier arrive/try_wait


@pytest.mark.parametrize("H", [12, 24, 48, 96, 192])
def test_cxx_tables_bit_exact_with_reference(golden_orders, H):
    assert np.array_equal(ops.order_table_host(_lib.ORDER_DILATION, H, 4), golden_orders[f"dilation_{H}_4"])
    assert np.array_equal(ops.order_table_host(_lib.ORDER_WINDOW, H, WS[H]), golden_orders[f"window_{H}_{WS[H]}"])
    line = ops.order_table_host(_lib.ORDER_LINE, H)
    assert np.array_equal(line[4:], golden_orders[f"line_{H}"])
    assert np.array_equal(line[:4], ops.order_table_host(_lib.ORDER_STD, H))


@pytest.mark.parametrize("H", [6, 7, 14, 28, 56])
def test_cxx_line_tables_other_sizes(golden_orders, H):
    assert np.array_equal(ops.order_table_host(_lib.ORDER_LINE, H)[4:], golden_orders[f"line_{H}"])


def test_cxx_tables_reference_main_blocks(golden_orders):
    assert np.array_equal(ops.order_table_host(_lib.ORDER_DILATION, 4, 3), golden_orders["dilation_4_3"])
    assert np.array_equal(ops.order_table_host(_lib.ORDER_WINDOW, 8, 4), golden_orders["window_8_4"])


def test_std_table_matches_oracle():
    from oracle import scan_orders as so
    for H in (3, 12, 17):
        assert np.array_equal(ops.order_table_host(_lib.ORDER_STD, H), so.std_indices(H, H))


@pytest.mark.parametrize("order,H,param", [(_lib.ORDER_LINE, 12, 0), (_lib.ORDER_LINE, 96, 0),
                                           (_lib.ORDER_WINDOW, 24, 8), (_lib.ORDER_DILATION, 48, 4)])
def test_plan_inverse_is_the_sorted_csr(order, H, param):
    K, plan = ops.order_plan_host(order, H, param)
    L = H * H
    fwd = plan[:K * L].reshape(K, L)
    ent = plan[K * L:2 * K * L].reshape(K, L)
    off = plan[2 * K * L:].reshape(K, L + 1)
    for k in range(K):
        assert off[k, 0] == 0 and off[k, L] == L
        order_ = np.argsort(fwd[k], kind="stable")       # tokens sorted by (pixel, t)
        assert np.array_equal(ent[k], order_)
        assert np.array_equal(np.diff(off[k]), np.bincount(fwd[k], minlength=L))


@pytest.mark.parametrize("order,H,param", [(_lib.ORDER_STD, 12, 0), (_lib.ORDER_STD, 96, 0), (_lib.ORDER_LINE, 12, 0),
                                           (_lib.ORDER_LINE, 24, 0), (_lib.ORDER_LINE, 48, 0), (_lib.ORDER_LINE, 96, 0),
                                           (_lib.ORDER_WINDOW, 24, 8), (_lib.ORDER_WINDOW, 96, 16),
                                           (_lib.ORDER_DILATION, 48, 4), (_lib.ORDER_DILATION, 96, 4)])
def test_permutation_schedule_properties(order, H, param):
    """tramba_order_sched (what the fused backward walks): per direction the moves (pixel, token) of the order table,
    every token exactly once, nothing else; identity / reversed-identity / permutation / repeated-pixel kinds; a
    permutation uses L entries (rounded up to the granule); and every warp step (32 moves) touches 32 distinct pixel banks AND 32 distinct
    token banks -- the property that makes the shared-memory permutation bank-conflict-free in both directions."""
    K = _lib.lib.tramba_order_num_directions(order)
    L = H * H
    tab = ops.order_table_host(order, H, param).reshape(K, L)
    n = int(_lib.lib.tramba_order_sched_len(order, H, H, param))
    assert n > 0 and n % max(64, 2 * L // 18) == 0
    sched, kinds = np.empty((K, n), np.uint32), np.empty((2 * K,), np.int32)
    _lib.check(_lib.lib.tramba_order_sched(order, H, H, param, sched.ctypes.data, kinds.ctypes.data))
    for k in range(K):
        used = int(kinds[K + k])
        assert 0 < used <= n and (sched[k, used:] == 0xFFFFFFFF).all()
        e = sched[k, :used]
        v = e[e != 0xFFFFFFFF].astype(np.int64)
        tok, pix = (v >> 16) >> 1, (v & 0xFFFF) >> 1
        assert np.array_equal(np.sort(tok), np.arange(L)) and np.array_equal(tab[k][tok], pix)
        ident, rev = np.array_equal(tab[k], np.arange(L)), np.array_equal(tab[k], np.arange(L)[::-1])
        dup = len(np.unique(tab[k])) < L
        assert kinds[k] == (0 if ident else 1 if rev else 3 if dup else 2)
        if not dup:                                            # a permutation: L moves in the fewest whole granules
            g = max(64, 2 * L // 18)
            assert used == (L + g - 1) // g * g
        steps = e.reshape(-1, 32, 2)                           # 64-entry blocks = two interleaved steps of 32 lanes
        for c in range(2):
            for row in steps[:, :, c]:
                r = row[row != 0xFFFFFFFF].astype(np.int64)
                pb, tb = ((r & 0xFFFF) >> 1) & 31, ((r >> 16) >> 1) & 31
                assert len(np.unique(pb)) == len(pb) and len(np.unique(tb)) == len(tb)


def test_order_errors():
    buf = np.zeros(16, np.int32)
    assert _lib.lib.tramba_order_table(_lib.ORDER_WINDOW, 12, 12, 5, buf.ctypes.data) == -4   # 12 % 5
    assert _lib.lib.tramba_order_table(_lib.ORDER_STD, 12, 13, 0, buf.ctypes.data) == -2      # not square
    assert _lib.lib.tramba_order_table(99, 12, 12, 0, buf.ctypes.data) == -4
    assert _lib.lib.tramba_order_table(_lib.ORDER_STD, 2, 2, 0, None) == -1
    with pytest.raises(KeyError):
        ops.window_size_for["13"]                        # reference: window_index[str(H)] KeyError


def test_scan_merge_argument_checks_cover_the_ablation_orders():
    """Order / K validation happens before the device check, so it is observable without a GPU."""
    import torch
    a = np.zeros(256, np.float32)
    p = a.ctypes.data
    plan = np.zeros(4096, np.int32).ctypes.data
    scan, merge = _lib.lib.tramba_cross_scan, _lib.lib.tramba_cross_merge
    assert scan(p, p, plan, _lib.ORDER_SPIRAL, 4, 1, 1, 4, 4, _lib.F32, None) == -4      # spiral has K = 8
    assert merge(p, p, plan, _lib.ORDER_HILBERT, 8, 1, 1, 4, 4, _lib.F32, None) == -4    # hilbert has K = 4
    assert scan(p, p, plan, _lib.ORDER_AB1 + 1, 4, 1, 1, 4, 4, _lib.F32, None) == -4     # past the last family
    if not torch.cuda.is_available():
        assert scan(p, p, plan, _lib.ORDER_SPIRAL, 8, 1, 1, 4, 4, _lib.F32, None) == -6  # valid arguments: no device
        assert scan(p, p, plan, _lib.ORDER_CUSTOM, 3, 1, 1, 4, 4, _lib.F32, None) == -6  # any K with a caller table
        assert merge(p, p, plan, _lib.ORDER_DS, 8, 1, 1, 4, 4, _lib.F32, None) == -6
    # the fused core never accepts them (they run op-by-op)
    assert _lib.lib.tramba_ss2d_fused_supported(_lib.ORDER_SPIRAL, 8, 64, 24, 24, 1, _lib.BF16, 1) == 0
    assert _lib.lib.tramba_ss2d_fused_supported(_lib.ORDER_LINE, 8, 64, 24, 24, 1, _lib.BF16, 1) == 1


def test_dct_basis_bit_exact(golden_dct):
    for n in (12, 24, 48, 96):
        w = np.empty((n, n), np.float32)
        _lib.lib.tramba_dct_basis(n, w.ctypes.data)
        assert np.array_equal(w, golden_dct[f"dct_{n}_wx"])


def test_state_dict_surface():
    m = tramba_b200.SS2D(d_model=32, d_state=1, channel_first=True, k_group=8,
                         scan=tramba_b200.CrossScan_Line, merge=tramba_b200.CrossMerge_Line)
    sd = m.state_dict()
    assert sd["x_proj_weight"].shape == (8, 2 + 2, 64)
    assert sd["dt_projs_weight"].shape == (8, 64, 2) and sd["dt_projs_bias"].shape == (8, 64)
    assert sd["A_logs"].shape == (8 * 64, 1) and sd["Ds"].shape == (8 * 64,)
    assert set(sd) == {"x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds", "out_norm.weight",
                       "out_norm.bias", "in_proj.weight", "conv2d.weight", "out_proj.weight"}
    d = tramba_b200.DCT2D(24, 24)
    assert set(d.state_dict()) == {"dct_x.weight", "dct_y.weight"}


def test_no_cpu_path_and_loud_errors():
    import torch
    with pytest.raises(RuntimeError, match="no CPU path"):
        tramba_b200.cross_scan(torch.zeros(1, 1, 12, 12), _lib.ORDER_STD)
    with pytest.raises(RuntimeError):
        tramba_b200.selective_scan_cuda_oflex.fwd(torch.zeros(1, 4, 12), torch.zeros(1, 4, 12), torch.zeros(4, 1),
                                                  torch.zeros(1, 1, 1, 12), torch.zeros(1, 1, 1, 12), None, None,
                                                  True, 1, True)
    if not torch.cuda.is_available():
        # compute entry points refuse to run without an sm_100 device (host pointers, no launch)
        a = np.zeros(64, np.float32)
        rc = _lib.lib.tramba_dct2d_fwd(a.ctypes.data, a.ctypes.data, a.ctypes.data, a.ctypes.data,
                                       a.ctypes.data, 1, 4, 4, None)
        assert rc == -6 and b"sm_100" in _lib.lib.tramba_error_string(rc)


def test_product_never_imports_oracle_or_reference():
    pkg = os.path.join(ROOT, "tramba_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "/root/reference" not in src, f


def test_extension_module_alias():
    import sys
    mod = tramba_b200.install_extension()
    import selective_scan_cuda_oflex as ext
    assert ext is mod and callable(ext.fwd) and callable(ext.bwd)
    del sys.modules["selective_scan_cuda_oflex"]


def test_sscan_ckpt_elems_plan():
    f = _lib.lib.tramba_sscan_ckpt_elems
    assert f(4, 1024, 9216, 1, _lib.F32) == 4 * 1024 * 24          # TSEG 12: 24 chunks of 384 per row
    assert f(4, 4096, 576, 1, _lib.F32) == 4 * 4096 * 2            # TSEG 9: 2 chunks of 288 per row
    assert f(2, 8192, 144, 1, _lib.F32) == 2 * 8192 // 2           # two 144-token rows per 288 chunk
    assert f(1, 8, 50, 3, _lib.F32) > 0                            # generic path


def test_c_consumer_builds_and_runs(tmp_path):
    """examples/abi_demo.c: the header is valid (pedantic) C99 and a plain C program links the library and uses the
    host-side entry points -- no torch, no Python, no CUDA headers on the consumer side."""
    import shutil
    import torch
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "abi_demo")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "abi_demo.c"), "-L" + libdir, "-ltramba_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    env = dict(os.environ)
    if not torch.cuda.is_available():
        env["TRAMBA_DEMO_NO_DEVICE"] = "1"
    out = subprocess.run([exe], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "window K=4 first tokens 0 1 24 575" in out.stdout and "helix K=8 plan elems 13832" in out.stdout
    if not torch.cuda.is_available():
        assert "cross_scan without a device -> -6" in out.stdout


def test_checkpoint_surface_linear2d_and_ss2d_roundtrip():
    """SURVEY 5 (checkpoint / resume): reference checkpoints must load into the drop-ins.  Linear2d accepts both the
    nn.Linear-shaped (out, in) weight and a conv-shaped (out, in, 1, 1) one (Models/modules.py:15-19); an SS2D
    state_dict round-trips with strict=True and identical values."""
    import torch
    lin = tramba_b200.Linear2d(6, 4, bias=True)
    w = torch.randn(4, 6)
    lin.load_state_dict({"weight": w.view(4, 6, 1, 1), "bias": torch.zeros(4)}, strict=True)
    assert torch.equal(lin.weight, w) and lin.weight.shape == (4, 6)
    lin.load_state_dict({"weight": 2 * w, "bias": torch.ones(4)}, strict=True)
    assert torch.equal(lin.weight, 2 * w)
    a = tramba_b200.SS2D(d_model=16, d_state=1, channel_first=True, k_group=4)
    b = tramba_b200.SS2D(d_model=16, d_state=1, channel_first=True, k_group=4)
    missing = b.load_state_dict(a.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    for (k, va), (_, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(va, vb), k
