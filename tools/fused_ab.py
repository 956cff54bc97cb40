#!/usr/bin/env python
"""A/B of the two fused SS2D backward paths on the GPU, per Tramba-V call site, batch 4:
   v1 = tramba_ss2d_fused_bwd (du in scan order to HBM) + tramba_cross_merge_skewed
   dx = tramba_ss2d_fused_bwd_dx (merge folded in, dx fp32) + the cast of dx to the operand dtype
CUDA events, L2 flushed between repetitions; prints one JSON line per site and the per-step total weighted by calls.
    python tools/fused_ab.py [--dtype bf16] [--reps 20] [--only enc0]"""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tramba_b200 import _lib, fused

SITES = [  # name, order, param, K, d_inner, H, calls per image
    ("enc0/std", _lib.ORDER_STD, 0, 4, 256, 96, 2), ("enc1/std", _lib.ORDER_STD, 0, 4, 512, 48, 2),
    ("enc2/std", _lib.ORDER_STD, 0, 4, 1024, 24, 15), ("enc3/std", _lib.ORDER_STD, 0, 4, 2048, 12, 2),
    ("freq96/win", _lib.ORDER_WINDOW, 16, 4, 256, 96, 1), ("freq96/dil", _lib.ORDER_DILATION, 4, 4, 256, 96, 1),
    ("freq48/win", _lib.ORDER_WINDOW, 12, 4, 512, 48, 1), ("freq48/dil", _lib.ORDER_DILATION, 4, 4, 512, 48, 1),
    ("freq24/win", _lib.ORDER_WINDOW, 8, 4, 1024, 24, 1), ("freq24/dil", _lib.ORDER_DILATION, 4, 4, 1024, 24, 1),
    ("dec96/line", _lib.ORDER_LINE, 0, 8, 256, 96, 2), ("dec48/line", _lib.ORDER_LINE, 0, 8, 512, 48, 2),
    ("dec24/line", _lib.ORDER_LINE, 0, 8, 1024, 24, 2),
]


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dt = torch.float32 if a.dtype == "f32" else torch.bfloat16
    code = _lib.F32 if a.dtype == "f32" else _lib.BF16
    es = 4 if a.dtype == "f32" else 2
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    tot = {"v1": 0.0, "dx": 0.0}
    for name, order, param, K, D, H, calls in SITES:
        if a.only and a.only not in name:
            continue
        B, L = a.batch, H * H
        x = torch.randn(B, D, H, H, device=dev).to(dt)
        dts = torch.randn(B, K, D, L, device=dev).to(dt)
        Bs, Cs = torch.randn(B, K, L, device=dev).to(dt), torch.randn(B, K, L, device=dev).to(dt)
        A = -(0.5 + torch.rand(K * D, device=dev))
        Dp = torch.ones(K * D, device=dev)
        bias = torch.randn(K * D, device=dev) - 3
        dy = torch.randn(B, D, L, device=dev)
        ddt = torch.empty_like(dts)
        dBs, dCs = torch.empty(B, K, L, device=dev), torch.empty(B, K, L, device=dev)
        pA, pD, pb = (torch.empty(B, K * D, device=dev) for _ in range(3))
        r = dict(site=name, dtype=a.dtype, tokens=K * B * D * L)
        nbytes = K * B * D * L * (4 * es + 4) + 2 * B * K * L * (es + 4)
        if _lib.lib.tramba_ss2d_fused_supported(order, K, D, H, H, 1, code, 1):
            idx = fused._idx16(order, H, param, dev)
            du = torch.empty_like(dts)
            dx1 = torch.empty(B, D, L, device=dev, dtype=dt)

            def v1():
                _lib.check(_lib.lib.tramba_ss2d_fused_bwd(
                    x.data_ptr(), dts.data_ptr(), Bs.data_ptr(), Cs.data_ptr(), A.data_ptr(), Dp.data_ptr(),
                    bias.data_ptr(), idx.data_ptr(), dy.data_ptr(), du.data_ptr(), ddt.data_ptr(), dBs.data_ptr(),
                    dCs.data_ptr(), pA.data_ptr(), pD.data_ptr(), pb.data_ptr(), order, K, B, D, H, H, 1, code, 1, st()))
                _lib.check(_lib.lib.tramba_cross_merge_skewed(du.data_ptr(), dx1.data_ptr(), idx.data_ptr(), order, K,
                                                              B, D, H, H, code, st()))
            r["v1_us"] = round(timeit(v1, a.reps, flush), 1)
            tot["v1"] += calls * r["v1_us"]
            dd1 = ddt.clone()
        if _lib.lib.tramba_ss2d_fused_bwd_dx_supported(order, K, D, H, H, 1, code):
            sched, kinds = fused._sched(order, H, param, dev)
            dx32 = torch.empty(B, D, L, device=dev)

            def v2(cast=True):
                _lib.check(_lib.lib.tramba_ss2d_fused_bwd_dx(
                    x.data_ptr(), dts.data_ptr(), Bs.data_ptr(), Cs.data_ptr(), A.data_ptr(), Dp.data_ptr(),
                    bias.data_ptr(), sched.data_ptr(), sched.shape[1], kinds.ctypes.data, dy.data_ptr(),
                    dx32.data_ptr(), ddt.data_ptr(), dBs.data_ptr(), dCs.data_ptr(), pA.data_ptr(), pD.data_ptr(),
                    pb.data_ptr(), order, K, B, D, H, H, 1, code, 1, st()))
                return dx32.to(dt) if cast else None
            r["dx_us"] = round(timeit(v2, a.reps, flush), 1)
            r["dx_kernel_us"] = round(timeit(lambda: v2(False), a.reps, flush), 1)
            r["dx_gbs_8d"] = round(nbytes / r["dx_kernel_us"] / 1e3, 1)
            tot["dx"] += calls * r["dx_us"]
            if "v1_us" in r:   # agreement of the two paths (different rounding points: bf16 du vs fp32 dx)
                v1()
                v2()
                torch.cuda.synchronize()
                r["ddelta_max_abs_diff"] = float((dd1.float() - ddt.float()).abs().max())
                r["dx_rel_diff"] = float((dx1.float() - dx32).abs().max() / dx32.abs().max())
        # ---- forward: first generation (skewed planes, table gathers) vs paired directions on the schedule
        y1 = torch.empty(B, D, L, device=dev)
        if _lib.lib.tramba_ss2d_fused_supported(order, K, D, H, H, 1, code, 0):
            idx = fused._idx16(order, H, param, dev)
            f1 = lambda: _lib.check(_lib.lib.tramba_ss2d_fused_fwd(
                x.data_ptr(), dts.data_ptr(), Bs.data_ptr(), Cs.data_ptr(), A.data_ptr(), Dp.data_ptr(), bias.data_ptr(),
                idx.data_ptr(), y1.data_ptr(), order, K, B, D, H, H, 1, code, 1, st()))
            r["fwd_v1_us"] = round(timeit(f1, a.reps, flush), 1)
            tot["fwd_v1"] = tot.get("fwd_v1", 0.0) + calls * r["fwd_v1_us"]
        if _lib.lib.tramba_ss2d_fused_fwd_sched_supported(order, K, D, H, H, 1, code):
            sched, kinds = fused._sched(order, H, param, dev)
            partners = fused._pairs(order, H, param)
            y2 = torch.empty(B, D, L, device=dev)
            f2 = lambda: _lib.check(_lib.lib.tramba_ss2d_fused_fwd_sched(
                x.data_ptr(), dts.data_ptr(), Bs.data_ptr(), Cs.data_ptr(), A.data_ptr(), Dp.data_ptr(), bias.data_ptr(),
                sched.data_ptr(), sched.shape[1], kinds.ctypes.data, partners.ctypes.data,
                fused._idx16_plain(order, H, param, dev).data_ptr(), y2.data_ptr(), order, K, B, D, H, H, 1, code, 1, st()))
            r["fwd_pair_us"] = round(timeit(f2, a.reps, flush), 1)
            fb = (B * D * L) * (es + 4) + K * B * D * L * es + 2 * B * K * L * es
            r["fwd_pair_gbs_fused"] = round(fb / r["fwd_pair_us"] / 1e3, 1)
            tot["fwd_pair"] = tot.get("fwd_pair", 0.0) + calls * r["fwd_pair_us"]
            if "fwd_v1_us" in r:
                f1(), f2()
                torch.cuda.synchronize()
                r["y_rel_diff"] = float((y1 - y2).abs().max() / y1.abs().max())
        print(json.dumps(r), flush=True)
    print(json.dumps({"us_per_image_step_weighted_by_calls": {k: round(v, 1) for k, v in tot.items()}}), flush=True)


if __name__ == "__main__":
    main()
