"""BASELINE.json configs 4 and 5 on one GPU: Tramba-V forward (eval, no_grad, bf16 autocast) images/s.
  config 4: 384x384, batch 32 (inference shard; N GPUs = N independent replicas of this)
  config 5: 768x768, batch 8 (stress: 192x192 maps -- build-side tables: dilation 4, window 32)
The forward is captured in a CUDA graph (static shapes); CUDA-event timed, inputs larger than L2.
Usage: python tools/infer_bench.py [--configs 4,5] [--steps 10]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tramba_b200  # noqa: F401
from tramba_b200.model import build_model


def run(img, batch, steps, deep):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = build_model(deep_supervision=deep, img_size=img, dims=128).to(dev).eval()
    x = torch.randn(batch, 3, img, img, device=dev)

    def fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(x)

    for _ in range(3):
        out = fwd()
    torch.cuda.synchronize()
    graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fwd()
        g.replay()
        torch.cuda.synchronize()
        graph = g
    except Exception as exc:  # report and fall back to eager launches
        print(f"graph capture failed: {type(exc).__name__}: {str(exc)[:200]}", file=sys.stderr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        graph.replay() if graph is not None else fwd()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    o = out[-1] if isinstance(out, (list, tuple)) else out
    return {"img": img, "batch": batch, "ms_per_forward": round(ms, 3), "images_per_s": round(batch / ms * 1e3, 1),
            "cuda_graph": graph is not None, "out_shape": list(o.shape), "finite": bool(torch.isfinite(o.float()).all())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="4,5")
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    for c in a.configs.split(","):
        if c == "4":
            r = run(384, 32, a.steps, deep=False)
            r["config"] = "4: Tramba-V-SOD inference 384^2, batch 32 per GPU"
        else:
            r = run(768, 8, a.steps, deep=False)
            r["config"] = "5: Tramba-V forward 768^2, batch 8 (192x192 maps)"
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
