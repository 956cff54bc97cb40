"""BASELINE.json config 2: single SS2D / Helix-SS2D block forward + backward microbenchmark --
batch 4, 96x96 tokens, d_model 128 (d_inner 256), K = 4 (CrossScan) and K = 8 (CrossScan_Line), d_state 16 (the
config's value; op-by-op kernels) and d_state 1 (the model's value; fused core).  fp32 and bf16 autocast;
CUDA-event timed over `--reps` iterations after warm-up, inputs 4 x 128 x 96 x 96.
Usage: python tools/ss2d_bench.py [--reps 20]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import tramba_b200
from tramba_b200 import ops


def run(K, N, amp, reps):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    scan, merge = (ops.CrossScan, ops.CrossMerge) if K == 4 else (ops.CrossScan_Line, ops.CrossMerge_Line)
    m = tramba_b200.SS2D(d_model=128, d_state=N, ssm_ratio=2.0, dt_rank="auto", act_layer=nn.SiLU, d_conv=3,
                         conv_bias=False, channel_first=True, disable_z=True, scan=scan, merge=merge,
                         k_group=K).to(dev)
    x = torch.randn(4, 128, 96, 96, device=dev, requires_grad=True)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            y = m(x)
        y.float().sum().backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tokens = 4 * K * 256 * 96 * 96
    return {"block": "SS2D" if K == 4 else "Helix-SS2D", "K": K, "d_state": N, "dtype": "bf16" if amp else "f32",
            "ms_fwd_bwd": round(ms, 3), "gtok_per_s": round(tokens / ms / 1e6, 1),
            "path": "fused core" if N == 1 and (amp or K == 4) else "see tramba_ss2d_fused_supported"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    for K in (4, 8):
        for N in (16, 1):
            for amp in (False, True):
                print(json.dumps(run(K, N, amp, a.reps)), flush=True)


if __name__ == "__main__":
    main()
