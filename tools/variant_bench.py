#!/usr/bin/env python
"""A/B of the fused SS2D core kernels of two builds of the library on the GPU, in ONE process:

    python -m tramba_b200.build --next          # (CPU box) libtramba_b200_next.so = candidates under csrc/next/
    python tools/variant_bench.py [--reps 20] [--dtype bf16] [--only enc0]

For every SS2D call site of Tramba-V at 384^2, batch 4: runs tramba_ss2d_fused_fwd / _bwd of the
product library and of the candidate library on the SAME device buffers, checks the candidate against the product
(y, du, ddelta, dA / dD / dbias partials bit-identical -- the candidates keep the floating-point operation order;
dB / dC go through global fp32 reds, so allclose) and times both (CUDA events, L2 flushed between reps).
Prints one JSON line per site and a summary weighted by calls per image.  The product library is the one whose
results are GPU-verified against the oracle (tests/test_gpu_fused.py), so agreement with it is parity by transit.
"""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tramba_b200 import _lib

SITES = [  # name, order, param, K, d_inner, H, calls per image
    ("enc0/std", _lib.ORDER_STD, 0, 4, 256, 96, 2), ("enc1/std", _lib.ORDER_STD, 0, 4, 512, 48, 2),
    ("enc2/std", _lib.ORDER_STD, 0, 4, 1024, 24, 15), ("enc3/std", _lib.ORDER_STD, 0, 4, 2048, 12, 2),
    ("freq96/win", _lib.ORDER_WINDOW, 16, 4, 256, 96, 1), ("freq96/dil", _lib.ORDER_DILATION, 4, 4, 256, 96, 1),
    ("freq48/win", _lib.ORDER_WINDOW, 12, 4, 512, 48, 1), ("freq48/dil", _lib.ORDER_DILATION, 4, 4, 512, 48, 1),
    ("freq24/win", _lib.ORDER_WINDOW, 8, 4, 1024, 24, 1), ("freq24/dil", _lib.ORDER_DILATION, 4, 4, 1024, 24, 1),
    ("dec96/line", _lib.ORDER_LINE, 0, 8, 256, 96, 2), ("dec48/line", _lib.ORDER_LINE, 0, 8, 512, 48, 2),
    ("dec24/line", _lib.ORDER_LINE, 0, 8, 1024, 24, 2),
]


def load(path):
    lib = ctypes.CDLL(path)
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def check(lib, rc):
    if rc != 0:
        raise RuntimeError(lib.tramba_error_string(rc).decode())


def timeit(fn, reps, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(200000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3   # us


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--base", default=os.path.join(os.path.dirname(_lib.LIB_PATH), "libtramba_b200.so"))
    ap.add_argument("--cand", default=os.path.join(os.path.dirname(_lib.LIB_PATH), "libtramba_b200_next.so"))
    a = ap.parse_args(argv)
    libs = {"base": load(a.base), "cand": load(a.cand)}
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    code = _lib.BF16 if a.dtype == "bf16" else _lib.F32
    es = 2 if a.dtype == "bf16" else 4
    dev = torch.device(a.device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream if dev.type == "cuda" else 0)
    B = a.batch
    tot = {"base": [0.0, 0.0, 0.0], "cand": [0.0, 0.0, 0.0]}
    ok_all = True
    for name, order, param, K, D, H, calls in SITES:
        if a.only and a.only not in name:
            continue
        L = H * H
        if not libs["base"].tramba_ss2d_fused_supported(order, K, D, H, H, 1, code, 1):
            print(json.dumps({"site": name, "skipped": "fused core does not cover this shape/dtype"}))
            continue
        g = torch.Generator(device="cpu").manual_seed(H + K)
        x = torch.randn(B, D, L, generator=g).to(dev, dt)
        dts = (0.5 * torch.randn(B, K, D, L, generator=g)).to(dev, dt)
        Bs = torch.randn(B, K, L, generator=g).to(dev, dt)
        Cs = torch.randn(B, K, L, generator=g).to(dev, dt)
        A = (-torch.rand(K * D, generator=g) - 0.5).to(dev)
        Dp = torch.ones(K * D, device=dev)
        bias = (torch.rand(K * D, generator=g) - 3.0).to(dev)
        dy = torch.randn(B, D, L, generator=g).to(dev)
        t16 = np.empty((K, L), dtype=np.uint16)
        check(libs["base"], libs["base"].tramba_order_table16_skewed(order, H, H, param, t16.ctypes.data))
        idx = torch.from_numpy(t16.view(np.int16)).to(dev)
        outs, times = {}, {}
        for tag, lib in libs.items():
            y = torch.empty(B, D, L, device=dev)
            du = torch.empty(B, K, D, L, device=dev, dtype=dt)
            dd = torch.empty(B, K, D, L, device=dev, dtype=dt)
            dB = torch.empty(B, K, L, device=dev)
            dC = torch.empty(B, K, L, device=dev)
            part = torch.empty(3, B, K * D, device=dev)

            def fwd(lib=lib, y=y):
                check(lib, lib.tramba_ss2d_fused_fwd(x.data_ptr(), dts.data_ptr(), Bs.data_ptr(), Cs.data_ptr(),
                                                     A.data_ptr(), Dp.data_ptr(), bias.data_ptr(), idx.data_ptr(),
                                                     y.data_ptr(), order, K, B, D, H, H, 1, code, 1, st))

            def bwd(lib=lib, du=du, dd=dd, dB=dB, dC=dC, part=part):
                check(lib, lib.tramba_ss2d_fused_bwd(x.data_ptr(), dts.data_ptr(), Bs.data_ptr(), Cs.data_ptr(),
                                                     A.data_ptr(), Dp.data_ptr(), bias.data_ptr(), idx.data_ptr(),
                                                     dy.data_ptr(), du.data_ptr(), dd.data_ptr(), dB.data_ptr(),
                                                     dC.data_ptr(), part[0].data_ptr(), part[1].data_ptr(),
                                                     part[2].data_ptr(), order, K, B, D, H, H, 1, code, 1, st))

            dx = torch.empty(B, D, L, device=dev, dtype=dt)

            def merge(lib=lib, du=du, dx=dx):
                check(lib, lib.tramba_cross_merge_skewed(du.data_ptr(), dx.data_ptr(), idx.data_ptr(), order, K, B, D,
                                                         H, H, code, st))

            times[tag] = (timeit(fwd, a.reps, flush), timeit(bwd, a.reps, flush), timeit(merge, a.reps, flush))
            outs[tag] = (y, du, dd, part, dB, dC, dx)
        b_, c_ = outs["base"], outs["cand"]
        line_dup = order == _lib.ORDER_LINE    # helix directions merge duplicate pixels with smem atomics (fwd y)
        exact = {"y": torch.equal(b_[0], c_[0]) if not line_dup else bool(torch.allclose(b_[0], c_[0], rtol=1e-5, atol=1e-5)),
                 "du": torch.equal(b_[1], c_[1]), "ddelta": torch.equal(b_[2], c_[2]),
                 "dA_dD_dbias": torch.equal(b_[3], c_[3]),
                 "dB": bool(torch.allclose(b_[4], c_[4], rtol=1e-4, atol=1e-4 * float(b_[4].abs().max()))),
                 "dC": bool(torch.allclose(b_[5], c_[5], rtol=1e-4, atol=1e-4 * float(b_[5].abs().max())))}
        exact["dx_merge"] = torch.equal(b_[6], c_[6]) if not line_dup else bool(
            torch.allclose(b_[6].float(), c_[6].float(), rtol=1e-2, atol=1e-2 * float(b_[6].float().abs().max())))
        # not bit-identical (e.g. the compiler contracted an FMA differently): report how far, accept fp32 rounding
        drift = {}
        for key, i in (("y", 0), ("du", 1), ("ddelta", 2), ("dA_dD_dbias", 3)):
            if not exact[key]:
                d = (b_[i].float() - c_[i].float()).abs().max().item() / max(b_[i].float().abs().max().item(), 1e-30)
                drift[key] = d
                exact[key] = d <= (1e-2 if (a.dtype == "bf16" and key in ("du", "ddelta")) else 1e-5)
        ok = all(exact.values())
        ok_all &= ok
        fb = (B * D * L) * (es + 4) + K * B * D * L * es + 2 * B * K * L * es
        bb = K * B * D * L * (4 * es + 4) + 2 * B * K * L * (es + 4)
        for tag in libs:
            tot[tag][0] += calls * times[tag][0]
            tot[tag][1] += calls * times[tag][1]
            tot[tag][2] += calls * times[tag][2]
        print(json.dumps({"site": name, "dtype": a.dtype, "parity_vs_product": exact, "max_rel_drift_where_not_bit_identical": drift, "ok": ok,
                          "fwd_us": {k: round(v[0], 1) for k, v in times.items()},
                          "bwd_us": {k: round(v[1], 1) for k, v in times.items()},
                          "merge_us": {k: round(v[2], 1) for k, v in times.items()},
                          "fwd_gbs": {k: round(fb / v[0] / 1e3, 1) for k, v in times.items()},
                          "bwd_gbs": {k: round(bb / v[1] / 1e3, 1) for k, v in times.items()}}), flush=True)
    print(json.dumps({"summary_us_per_step_weighted_by_calls": {k: {"fwd": round(v[0], 1), "bwd": round(v[1], 1), "merge": round(v[2], 1)}
                                                                for k, v in tot.items()},
                      "all_parity_ok": ok_all}))
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
