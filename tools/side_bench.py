"""Microbenchmark of the ops either side of the scan core at Tramba-V's shapes:
LayerNorm2d(+GELU) and depthwise conv(+SiLU), forward and backward, through the C-ABI; CUDA-event timed,
L2 flushed between repetitions.  Prints algorithmic GB/s (and GFMA/s for the convs).
Usage: python tools/side_bench.py [--batch 4] [--reps 20]"""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tramba_b200 import _lib

DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


FLUSH_READ = False


def timeit(fn, reps, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        if FLUSH_READ:
            flush.view(torch.int32).max()   # evict by READING 256 MB: the L2 is left cold but clean
        else:
            flush.zero_()                   # evict by writing: the L2 is left full of dirty lines
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
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--flush", default="write", choices=["write", "read"])
    ap.add_argument("--only", default="", help="substring of the site name")
    a = ap.parse_args()
    global FLUSH_READ
    FLUSH_READ = a.flush == "read"
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    B = a.batch
    f32, bf = torch.float32, torch.bfloat16
    LN = [("null", 8, 8, f32, f32, 0), ("norm enc0", 128, 96, f32, bf, 0), ("out_norm enc0", 256, 96, f32, bf, 1), ("norm enc1", 256, 48, f32, bf, 0),
          ("out_norm enc1", 512, 48, f32, bf, 1), ("norm enc2", 512, 24, f32, bf, 0), ("out_norm enc2", 1024, 24, f32, bf, 1),
          ("out_norm enc3", 2048, 12, f32, bf, 1), ("patch_embed", 64, 192, bf, f32, 0), ("final expand", 128, 384, bf, f32, 0)]
    for name, C, H, idt, odt, act in LN:
        if a.only and a.only not in name:
            continue
        L = H * H
        x = torch.randn(B, C, L, device=dev).to(idt)
        w, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        y = torch.empty(B, C, L, device=dev, dtype=odt)
        mean, rstd = torch.empty(B, L, device=dev), torch.empty(B, L, device=dev)
        tiles = _lib.lib.tramba_layernorm2d_tiles(B, C, L)
        gw, gb = torch.empty(B * tiles, C, device=dev), torch.empty(B * tiles, C, device=dev)
        gy, gx = torch.randn_like(y), torch.empty_like(x)
        fwd = lambda: _lib.check(_lib.lib.tramba_layernorm2d_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(),
                                                                 mean.data_ptr(), rstd.data_ptr(), B, C, L, 1e-5, act,
                                                                 DT[idt], DT[odt], st()))
        bwd = lambda: _lib.check(_lib.lib.tramba_layernorm2d_bwd(gy.data_ptr(), x.data_ptr(), w.data_ptr(), b.data_ptr(),
                                                                 mean.data_ptr(), rstd.data_ptr(), gx.data_ptr(),
                                                                 gw.data_ptr(), gb.data_ptr(), B, C, L, act, DT[idt],
                                                                 DT[odt], st()))
        tf, tb = timeit(fwd, a.reps, flush), timeit(bwd, a.reps, flush)
        n = B * C * L
        byf, byb = n * (x.element_size() + y.element_size()), n * (2 * x.element_size() + y.element_size())
        print(json.dumps(dict(op="layernorm2d", site=name, C=C, H=H, act=act, fwd_us=round(tf * 1e6, 1),
                              fwd_gbs=round(byf / tf / 1e9), bwd_us=round(tb * 1e6, 1), bwd_gbs=round(byb / tb / 1e9))),
              flush=True)
    DW = [("conv2d enc0", 256, 96, 3, 2), ("conv2d enc1", 512, 48, 3, 2), ("conv2d enc2", 1024, 24, 3, 2),
          ("conv2d enc3", 2048, 12, 3, 2), ("dwms dec96", 512, 96, 7, 0), ("dwms dec48", 1024, 48, 7, 0),
          ("dwms dec24", 2048, 24, 7, 0)]
    for name, C, H, k, act in DW:
        if a.only and a.only not in name:
            continue
        x = torch.randn(B, C, H, H, device=dev).to(bf)
        w, b = torch.randn(C, 1, k, k, device=dev) / k, torch.zeros(C, device=dev)
        y, dy, dx = torch.empty_like(x), torch.randn_like(x), torch.empty_like(x)
        part = torch.empty(B, C, k * k + 1, device=dev)
        fwd = lambda: _lib.check(_lib.lib.tramba_dwconv2d_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), B, C,
                                                              H, H, k, act, _lib.BF16, st()))
        bwd = lambda: _lib.check(_lib.lib.tramba_dwconv2d_bwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), dy.data_ptr(),
                                                              dx.data_ptr(), part.data_ptr(), B, C, H, H, k, act,
                                                              _lib.BF16, st()))
        tf, tb = timeit(fwd, a.reps, flush), timeit(bwd, a.reps, flush)
        n = B * C * H * H
        print(json.dumps(dict(op="dwconv2d", site=name, C=C, H=H, k=k, act=act, fwd_us=round(tf * 1e6, 1),
                              fwd_gbs=round(n * 4 / tf / 1e9), fwd_gfma=round(n * k * k / tf / 1e9),
                              bwd_us=round(tb * 1e6, 1), bwd_gbs=round(n * 6 / tb / 1e9),
                              bwd_gfma=round(n * k * k * (3 if act else 2) / tb / 1e9))), flush=True)


if __name__ == "__main__":
    main()
