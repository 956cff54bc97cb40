"""Per-kernel microbenchmark at Tramba-V's call-site shapes: achieved algorithmic
GB/s of cross-scan, selective-scan fwd/bwd, cross-merge and DCT2D, CUDA-event timed, L2 flushed
between repetitions.  Usage: python tools/kernel_bench.py [--batch 4] [--dtype f32|bf16] [--reps 20]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tramba_b200
from tramba_b200 import _lib

SITES = [  # name, order, K, d_inner, H
    ("enc0/std", _lib.ORDER_STD, 4, 256, 96), ("enc1/std", _lib.ORDER_STD, 4, 512, 48),
    ("enc2/std", _lib.ORDER_STD, 4, 1024, 24), ("enc3/std", _lib.ORDER_STD, 4, 2048, 12),
    ("freq96/win", _lib.ORDER_WINDOW, 4, 256, 96), ("freq96/dil", _lib.ORDER_DILATION, 4, 256, 96),
    ("freq48/win", _lib.ORDER_WINDOW, 4, 512, 48), ("freq48/dil", _lib.ORDER_DILATION, 4, 512, 48),
    ("freq24/win", _lib.ORDER_WINDOW, 4, 1024, 24), ("freq24/dil", _lib.ORDER_DILATION, 4, 1024, 24),
    ("dec96/line", _lib.ORDER_LINE, 8, 256, 96),
    ("dec48/line", _lib.ORDER_LINE, 8, 512, 48), ("dec24/line", _lib.ORDER_LINE, 8, 1024, 24),
]
PARAM = {(_lib.ORDER_WINDOW, 96): 16, (_lib.ORDER_WINDOW, 48): 12, (_lib.ORDER_WINDOW, 24): 8,
         (_lib.ORDER_WINDOW, 12): 4}


def timeit(fn, reps, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(200000)   # keep the GPU busy while the CPU enqueues: no launch gap inside the events
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--nstate", type=int, default=1)
    ap.add_argument("--only", default="")
    ap.add_argument("--fused-only", action="store_true", help="time only the fused SS2D core kernels")
    a = ap.parse_args()
    dt = torch.float32 if a.dtype == "f32" else torch.bfloat16
    es = 4 if a.dtype == "f32" else 2
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ext = tramba_b200.selective_scan_cuda_oflex
    N = a.nstate
    rows = []
    for name, order, K, D, H in SITES:
        if a.only and a.only not in name:
            continue
        B, L = a.batch, H * H
        param = PARAM.get((order, H), 4 if order == _lib.ORDER_DILATION else 0)
        x = torch.randn(B, D, H, H, device=dev, dtype=dt)
        delta = torch.randn(B, K * D, L, device=dev, dtype=dt)
        A = -(0.5 + torch.rand(K * D, N, device=dev))
        Bm = torch.randn(B, K, N, L, device=dev, dtype=dt)
        Cm = torch.randn(B, K, N, L, device=dev, dtype=dt)
        Dp = torch.ones(K * D, device=dev)
        bias = torch.randn(K * D, device=dev) - 3
        r = dict(site=name, B=B, K=K, D=D, H=H, dtype=a.dtype, N=N)
        from tramba_b200 import fused
        if N == 1 and fused.supported(order, K, D, H, H, 1, dt, need_bwd=False):
            dts_img = delta.view(B, K, D, L)
            Bi, Ci = Bm.view(B, K, L), Cm.view(B, K, L)
            Af = A.view(-1)
            idx = fused._idx16(order, H, param, dev)
            y = torch.empty(B, D, L, device=dev)
            code = _lib.F32 if a.dtype == "f32" else _lib.BF16
            import ctypes
            st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            ffwd = lambda: _lib.check(_lib.lib.tramba_ss2d_fused_fwd(
                x.data_ptr(), dts_img.data_ptr(), Bi.data_ptr(), Ci.data_ptr(), Af.data_ptr(), Dp.data_ptr(),
                bias.data_ptr(), idx.data_ptr(), y.data_ptr(), order, K, B, D, H, H, 1, code, 1, st()))
            t_ff = timeit(ffwd, a.reps, flush)
            by_ff = B * D * L * (es + 4) + K * B * D * L * es + 2 * B * K * L * es
            r.update(ffwd_us=t_ff * 1e6, ffwd_gbs=by_ff / t_ff / 1e9, ffwd_gtok=K * B * D * L / t_ff / 1e9)
            if fused.supported(order, K, D, H, H, 1, dt, need_bwd=True):
                dy = torch.randn(B, D, L, device=dev)
                sched, kinds = fused._sched(order, H, param, dev)
                n_dx, n_bc, n_row = B * D * L, B * K * L, K * D
                ws = torch.empty(n_dx + 2 * n_bc + 3 * n_row, device=dev)          # one workspace -> one memset
                dx, dBs, dCs, pA, pD, pb = torch.split(ws, [n_dx, n_bc, n_bc, n_row, n_row, n_row])
                ddt = torch.empty_like(dts_img)
                fbwd = lambda: _lib.check(_lib.lib.tramba_ss2d_fused_bwd(
                    x.data_ptr(), dts_img.data_ptr(), Bi.data_ptr(), Ci.data_ptr(), Af.data_ptr(), Dp.data_ptr(),
                    bias.data_ptr(), sched.data_ptr(), sched.shape[1], kinds.ctypes.data, dy.data_ptr(), dx.data_ptr(),
                    ddt.data_ptr(), dBs.data_ptr(), dCs.data_ptr(), pA.data_ptr(), pD.data_ptr(), pb.data_ptr(), order, K,
                    B, D, H, H, 1, code, 1, st()))
                t_fb = timeit(fbwd, a.reps, flush)
                by_fb = K * B * D * L * (4 * es + 4) + 2 * B * K * L * (es + 4)      # what it replaces
                by_min = B * D * L * (es + 4 + 4) + K * B * D * L * 2 * es + 2 * B * K * L * (es + 4)   # what it must move
                r.update(fbwd_us=t_fb * 1e6, fbwd_gbs=by_fb / t_fb / 1e9, fbwd_gbs_min=by_min / t_fb / 1e9,
                         fbwd_gtok=K * B * D * L / t_fb / 1e9)
        if a.fused_only:
            print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)
            continue
        xs = tramba_b200.cross_scan(x, order, param)
        u = xs.view(B, K * D, L)
        out, ck = ext.fwd(u, delta, A, Bm, Cm, Dp, bias, True, 1, True)
        dout = torch.randn_like(out)
        E = B * D * L
        t_scan = timeit(lambda: tramba_b200.cross_scan(x, order, param), a.reps, flush)
        t_fwd = timeit(lambda: ext.fwd(u, delta, A, Bm, Cm, Dp, bias, True, 1, True), a.reps, flush)
        t_bwd = timeit(lambda: ext.bwd(u, delta, A, Bm, Cm, Dp, bias, dout, ck, True, 1), a.reps, flush)
        ys = out.view(B, K, D, L)
        t_merge = timeit(lambda: tramba_b200.cross_merge(ys, order, H, H, param), a.reps, flush)
        by_scan = (1 + K) * E * es
        by_fwd = K * E * (2 * es + 4) + 2 * B * K * N * L * es
        by_bwd = K * E * (2 * es + 4) + K * E * 2 * es + 2 * B * K * N * L * (es + 4)
        by_merge = (1 + K) * E * 4
        r.update(
                 scan_gbs=by_scan / t_scan / 1e9, fwd_gbs=by_fwd / t_fwd / 1e9, bwd_gbs=by_bwd / t_bwd / 1e9,
                 merge_gbs=by_merge / t_merge / 1e9,
                 scan_us=t_scan * 1e6, fwd_us=t_fwd * 1e6, bwd_us=t_bwd * 1e6, merge_us=t_merge * 1e6)
        rows.append(r)
        print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)
    for C, H in ((128, 96), (256, 48), (512, 24)):
        if a.only and "dct" not in a.only:
            continue
        m = tramba_b200.DCT2D(H, H).to(dev)
        x = torch.randn(a.batch, C, H, H, device=dev)
        t = timeit(lambda: m(x), a.reps, flush)
        by = a.batch * C * H * H * 6
        fl = 3 * a.batch * C * H * H * H
        print(json.dumps(dict(site=f"dct{H}", us=round(t * 1e6, 1), gbs=round(by / t / 1e9, 1),
                              tflops=round(fl / t / 1e12, 2))), flush=True)


if __name__ == "__main__":
    main()
