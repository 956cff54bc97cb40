#!/usr/bin/env python
"""Where the 58 % of the training step that is NOT our kernels goes, by ATen operator and by model module.

    python tools/op_profile.py [--steps 3] [--batch 4] > gpurun_out/op_profile.txt

Runs the same training step as bench.py (Tramba-V-TSOD 384^2, bf16 autocast, fwd + loss + bwd + fused Adam) in
EAGER mode under torch.profiler (CPU + CUDA activities), with one record_function range per top-level block of the
harness, and prints (1) self device time per ATen operator (forward ops and their autograd nodes), (2) device time
per module range (forward only: the backward of a module runs outside its range).  Device times are CUPTI kernel
durations, so the eager launch gaps do not inflate them; the totals match bench.py --kernel-table.
Not part of any timed region; a tool for choosing what to fuse next.
"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile, record_function

import tramba_b200  # noqa: F401
from tramba_b200.model import build_model, tsod_loss


def dev_time(ev, self_only=True):
    for name in (("self_device_time_total", "self_cuda_time_total") if self_only else
                 ("device_time_total", "cuda_time_total")):
        v = getattr(ev, name, None)
        if v is not None:
            return float(v)
    return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--depth", type=int, default=2, help="module nesting depth that gets a range")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = build_model(deep_supervision=True, img_size=384, dims=128).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    x = torch.randn(a.batch, 3, 384, 384, device=dev)
    y = (torch.rand(a.batch, 1, 384, 384, device=dev) > 0.5).float()

    # one range per module down to --depth levels below the root (e.g. vssm_encoder.layers.2)
    def wrap(mod, name):
        fwd = mod.forward

        def ranged(*args, **kw):
            with record_function("mod:" + name):
                return fwd(*args, **kw)
        mod.forward = ranged

    for name, mod in model.named_modules():
        if name and name.count(".") < a.depth:
            wrap(mod, f"{name} [{type(mod).__name__}]")

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = model(x)
        loss = tsod_loss(outs, y)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()

    ops = collections.defaultdict(lambda: [0, 0.0])
    mods = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.key_averages():
        if ev.key.startswith("mod:"):
            mods[ev.key[4:]] = [ev.count, dev_time(ev, self_only=False)]
        else:
            t = dev_time(ev)
            if t > 0:
                ops[ev.key] = [ev.count, t]
    tot = sum(v[1] for v in ops.values())
    print(f"{a.steps} eager steps, batch {a.batch}: {tot / a.steps / 1e3:.2f} ms of device time per step")
    print("\nself device time per operator (autograd nodes appear under their own names):")
    for k, (n, t) in sorted(ops.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"{100 * t / tot:6.2f}% {t / a.steps / 1e3:8.3f} ms/step {n // a.steps:6d} x  {k[:100]}")
    # who calls the glue operators: device time of copy_ / sum / add / mul ... grouped by the kernel they launch, the
    # enclosing autograd node (backward) or module range (forward) found by walking the CPU parents, and the shapes
    callers = collections.defaultdict(lambda: [0, 0.0])
    glue = ("aten::copy_", "aten::sum", "aten::add", "aten::add_", "aten::mul", "aten::fill_", "aten::zero_",
            "aten::gelu", "aten::cat", "aten::clone", "aten::contiguous", "aten::gelu_backward", "aten::mul_")

    def enclosing(ev):
        node, mod = None, None
        p = getattr(ev, "cpu_parent", None)
        while p is not None:
            if node is None and (p.name.startswith("autograd::engine::evaluate_function: ") or p.name.endswith("Backward")):
                node = p.name.replace("autograd::engine::evaluate_function: ", "bwd:")
            if mod is None and p.name.startswith("mod:"):
                mod = p.name[4:]
            if node is None and not p.name.startswith(("aten::", "mod:", "autograd::", "cuda", "Optimizer", "ProfilerStep")):
                node = "in:" + p.name                      # a custom autograd Function / Python range
            p = getattr(p, "cpu_parent", None)
        return f"{node or '-'} @ {mod or '-'}"

    def kern_kind(ev):
        ks = [k.name for k in getattr(ev, "kernels", [])]
        if not ks:
            return "-"
        k = ks[0]
        for tag in ("unrolled_elementwise_kernel", "vectorized_elementwise_kernel", "elementwise_kernel", "reduce_kernel",
                    "multi_tensor_apply", "Memcpy", "memcpy"):
            if tag in k:
                extra = "direct_copy" if "direct_copy" in k else ("bfloat16_copy" if "bfloat16_copy" in k else "")
                return (tag + " " + extra).strip()
        return k[:40]

    for ev in prof.events():
        if ev.name not in glue:
            continue
        t = dev_time(ev)
        if t <= 0:
            continue
        shapes = str(getattr(ev, "input_shapes", ""))[:60]
        c = callers[(ev.name, kern_kind(ev), enclosing(ev), shapes)]
        c[0] += 1
        c[1] += t
    print("\nglue operators by launched kernel, enclosing autograd node / module range and input shapes (self device time):")
    for (op, kk, where, shapes), (n, t) in sorted(callers.items(), key=lambda kv: -kv[1][1])[:90]:
        print(f"{t / a.steps / 1e3:8.3f} ms/step {n // a.steps:5d} x  {op:14s} {kk:42s} {where[:70]:70s} {shapes}")
    print("\nforward device time per module range (nested ranges overlap; backward is not attributed):")
    for k, (n, t) in sorted(mods.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{t / a.steps / 1e3:8.3f} ms/step {n // a.steps:4d} x  {k}")


if __name__ == "__main__":
    main()
