#!/usr/bin/env python
"""bench.py -- Tramba-V 384^2 images/s, forward+backward (BASELINE.json metric, config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (sm_100a kernels)
    python bench.py --impl reference [--gpus N] [--steps K] ...     the reference's CPU path (oracle port)

A step = one full training step of Tramba-V-TSOD (tramba_b200/model.py harness: the reference's
module tree with identical state_dict) at 384x384, batch 4 per GPU, bf16 autocast: forward,
deep-supervision loss (train.py:74-85), backward, NCCL gradient all-reduce (N > 1), Adam step.
Random-init weights, synthetic N(0,1) images and Bernoulli(0.5) masks.

  value    : images/s, whole job, inputs already resident in HBM, CUDA events, max over ranks
  e2e      : same step through the public API with HOST inputs: pinned host -> device copy of
             images + masks and a device -> host read of the loss inside the timed region
  roofline : the dominant kernel of the scan path in the step (fused selective-scan backward); per-launch
             device time inside the graph-replayed step from CUPTI activity records (CUDA events around every
             launch in an eager region when CUPTI is unavailable); achieved = sum(algorithmic bytes) /
             sum(launch time) against MEASURED_PEAKS.json hbm_gbs; traffic = ncu DRAM bytes per launch
             (profiles/roofline_traffic.json)
  cpu_baseline : the oracle port of the reference path (same harness, oracle operators, fp32, CPU)
             on this box's host cores; N = 1 only; bounded sample = 1 image forward+backward
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Tramba-V 384x384 images/sec fwd+bwd (training step)"
UNIT = "images/s"
IMG = 384
NCU_TRAFFIC_PATH = os.path.join(ROOT, "profiles", "roofline_traffic.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 3 = training step (the headline metric, default); 2 = single SS2D block "
                         "fwd+bwd microbench; 4 = inference 384^2 batch 32; 5 = forward 768^2 batch 8 (single GPU)")
    ap.add_argument("--d-state", type=int, default=16, help="--config 2: d_state of the block (BASELINE: 16; model: 1)")
    ap.add_argument("--batch", type=int, default=4, help="images per GPU per step (config 3: 4)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step in a CUDA graph")
    ap.add_argument("--unfused", action="store_true", help="op-by-op SS2D path instead of the fused core")
    ap.add_argument("--overlap", action="store_true", help="N > 1: bucketed all-reduces launched from inside backward on "
                    "a side stream instead of one all-reduce after backward (measured slower, see run_ours)")
    ap.add_argument("--allreduce-dtype", default="bf16", choices=["bf16", "f32"], help="N > 1: dtype of the gradient "
                    "all-reduce (bf16 halves the bytes on the wire; the optimizer still reads fp32)")
    ap.add_argument("--no-zero", action="store_true", help="N > 1: replicated optimizer step (all-reduce of every "
                    "gradient, full Adam on every rank) instead of the ZeRO-1 step (reduce-scatter, sharded Adam, "
                    "all-gather of the bf16 weights)")
    ap.add_argument("--force-zero", action="store_true", help="run the ZeRO-1 step at N = 1 too (one shard: a functional "
                    "check of the sharded step without a second GPU)")
    ap.add_argument("--buckets", type=int, default=4, help="gradient buckets of the overlapped all-reduce")
    ap.add_argument("--no-shadow", action="store_true", help="plain autocast (weights cast on every use) instead of "
                    "bf16 shadow weights refreshed after the optimizer step")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="wall budget of the reference arm")
    ap.add_argument("--kernel-table", default="", help="write a per-kernel device-time table of 3 graph-replayed "
                    "steps (torch.profiler / CUPTI; not part of the timed regions) to this file")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        for ts, ln in self.lines:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), smax.append(float(f[1])), pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(self.NAMES, f[3:7]):
                if v == "Active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU path = harness + oracle operators (fp32, CPU)
# ----------------------------------------------------------------------------------------------
def cpu_path_step_fn(batch):
    """Returns (step, cores, sample description).  The ONLY place bench.py touches oracle/."""
    import torch
    from oracle import build as obuild
    from oracle.torch_ops import oracleize
    from tramba_b200.model import build_model, tsod_loss

    obuild.build()
    torch.manual_seed(0)
    cores = torch.get_num_threads()
    model = oracleize(build_model(deep_supervision=True, img_size=IMG, dims=128), f32=True).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    x = torch.randn(batch, 3, IMG, IMG)
    y = (torch.rand(batch, 1, IMG, IMG) > 0.5).float()

    def step():
        opt.zero_grad(set_to_none=True)
        loss = tsod_loss(model(x), y)
        loss.backward()
        opt.step()
        return loss.item()

    sample = (f"{batch} image(s) of the same training step (fwd + loss + bwd + Adam) on CPU: harness with the "
              f"oracle operators (C/OpenMP fp32 selective scan, numpy scan/merge/DCT), torch CPU for the rest")
    return step, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # This arm is the reference's CPU path: it must not map our CUDA library (the driver records which .so files a
    # bench process loads) and it uses every host core even under torchrun (which exports OMP_NUM_THREADS=1).
    os.environ["TRAMBA_B200_NO_LIB"] = "1"
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(ncores)
    os.environ["MKL_NUM_THREADS"] = str(ncores)
    import torch
    torch.set_num_threads(ncores)
    step, cores, sample = cpu_path_step_fn(1)
    t_begin = time.time()
    t = time.perf_counter()
    step()                                   # warm-up (also builds the order tables)
    t_first = time.perf_counter() - t
    warm = 1
    while warm < args.warmup and (time.time() - t_begin) + 2 * t_first < 0.25 * args.cpu_budget_s:
        step()
        warm += 1
    left = args.cpu_budget_s - (time.time() - t_begin)
    steps = max(1, min(args.steps, int(left / max(t_first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = steps * 1 / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "requested_steps": args.steps, "requested_warmup": args.warmup,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Tramba-V-TSOD training step 384x384 (BASELINE config 3), CPU sample batch 1",
                   "img_size": IMG, "global_batch": 1, "parallelism": "cpu", "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is pure Python + an un-vendored CUDA extension: no oracle/_ref; this is the oracle port. "
                "steps/warmup are clipped to --cpu-budget-s of wall time (requested_* = what was asked)",
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def summarize_prof(prof):
    """(kind -> {launches, us, bytes, gbs}) from the event list recorded by tramba_b200._lib.timed."""
    agg = {}
    for kind, nbytes, e0, e1 in prof:
        a = agg.setdefault(kind, {"launches": 0, "ms": 0.0, "bytes": 0})
        a["launches"] += 1
        a["ms"] += e0.elapsed_time(e1)
        a["bytes"] += nbytes
    for a in agg.values():
        a["gbs"] = a["bytes"] / (a["ms"] * 1e-3) / 1e9 if a["ms"] > 0 else None
        a["us_per_launch"] = 1e3 * a["ms"] / a["launches"]
    return agg


# kernel-name fragment -> the `kind` tramba_b200._lib.timed records for it
KIND_OF = {"ss2d_bwd_kernel": "ss2d_fused_bwd", "ss2d_fwd_kernel": "ss2d_fused_fwd", "batch_sum": "batch_sum",
           "ln2d_fwd": "layernorm2d_fwd", "ln2d_bwd": "layernorm2d_bwd", "dct2d_fwd": "dct2d_fwd", "dct2d_bwd": "dct2d_bwd",
           "proj_gather": "proj_gather", "proj_scatter": "proj_scatter", "dwconv_fwd": "dwconv_fwd", "dwconv_bwd": "dwconv_bwd",
           "sscan_fwd": "sscan_fwd", "sscan_bwd": "sscan_bwd", "cross_scan_kernel": "cross_scan", "cross_merge": "cross_merge"}


def kind_of(name):
    """bench `kind` (as recorded by tramba_b200._lib.timed) of a CUPTI kernel name, or None."""
    import re
    if "tramba" not in name:
        return None
    m = re.search(r"dwconv_(fwd|bwd)_kernel<[^,]+, (\d)", name)
    if m:
        return f"dwconv{m.group(2)}_{m.group(1)}"
    return next((v for kk, v in KIND_OF.items() if kk in name), None)


def kernel_times(local_step, steps=3, path=""):
    """Device time per kernel name over `steps` graph-replayed training steps (CUPTI activity records: warm
    caches, back-to-back launches -- the complement of ncu's serialised cold-cache launch list).  Returns
    {kernel name: [launches per step, us per step]}; optionally writes the table to `path`.
    `local_step` must be free of collectives: this runs on rank 0 only."""
    import collections
    import re
    import torch
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            local_step()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if getattr(ev, "device_type", None) is None or "cuda" not in str(ev.device_type).lower():
            continue
        name = ev.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"^void ", "", name)[:110]
        dur = getattr(ev, "device_time", None)
        if dur is None:
            dur = getattr(ev, "cuda_time", 0.0)
        agg[name][0] += 1
        agg[name][1] += float(dur)
    tot = sum(v[1] for v in agg.values())
    out = {k: [v[0] / steps, v[1] / steps] for k, v in agg.items()}
    if not path:
        return out
    with open(path, "w") as f:
        f.write(f"{steps} graph-replayed steps: {tot / steps / 1e3:.3f} ms device time per step in "
                f"{sum(v[0] for v in agg.values()) // steps} kernels/memcpys per step\n")
        ours = sum(v[1] for k, v in agg.items() if "tramba" in k)
        f.write(f"tramba:: kernels: {100 * ours / max(tot, 1e-9):.2f}% of device time\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:80]:
            f.write(f"{100 * v[1] / tot:6.2f}% {v[1] / steps / 1e3:9.3f} ms/step {v[0] // steps:6d} x "
                    f"{v[1] / v[0]:9.1f} us  {k}\n")
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    import tramba_b200
    from tramba_b200 import _lib
    from tramba_b200.model import build_model, tsod_loss

    torch.backends.cudnn.benchmark = True               # train.py:29
    torch.manual_seed(1026 + rank)                      # train.py:284 random_seed(1026)
    model = build_model(deep_supervision=True, img_size=IMG, dims=128).to(dev).train()
    if args.unfused:
        for m in model.modules():
            if isinstance(m, tramba_b200.SS2D):
                m.fused = False
    enc = [p for n, p in model.named_parameters() if n.startswith("vssm_encoder")]
    dec = [p for n, p in model.named_parameters() if not n.startswith("vssm_encoder")]
    opt = torch.optim.Adam([{"params": enc, "lr": 1e-5}, {"params": dec, "lr": 1e-4}], fused=True,
                           capturable=True)                                                      # train.py:266-280
    graph_error = None
    # Data parallel over images (the path shards by image): replicas start identical
    # (broadcast from rank 0) and every gradient lives in ONE flat fp32 buffer that is all-reduced
    # (NCCL, AVG) once per step between backward and the Adam step -- 446 MB over NVLink/NVSwitch.
    from tramba_b200.dp import FlatGradients, ShadowParams, ShardedStep
    params = [p for p in model.parameters()]
    amp = torch.bfloat16 if args.dtype == "bf16" else None
    if world > 1:
        for p in params:
            dist.broadcast(p.data, 0)
    # bf16 shadow copies of the conv / linear / projection weights (what autocast would cast on every use)
    shadow = ShadowParams(model, amp) if (amp is not None and not args.no_shadow) else None
    # N > 1 (default): ZeRO-1 for the shadowed weights (> 99 % of the parameters) -- reduce-scatter of the bf16 gradients,
    # Adam on this rank's 1/N shard of the fp32 masters, all-gather of the bf16 shadows; the fp32-only parameters stay
    # replicated (flat fp32 all-reduce + Adam), as everything does at N = 1.
    zs = None
    if (world > 1 or args.force_zero) and shadow is not None and not args.no_zero and not args.overlap:
        zs = ShardedStep(shadow, [(enc, 1e-5), (dec, 1e-4)], world, rank, fused=True, capturable=True)
        has = {id(m) for m in shadow.masters}
        params = [p for p in params if id(p) not in has]
        enc_r, dec_r = [p for p in enc if id(p) not in has], [p for p in dec if id(p) not in has]
        opt = torch.optim.Adam([{"params": enc_r, "lr": 1e-5}, {"params": dec_r, "lr": 1e-4}], fused=True, capturable=True)
        fg = FlatGradients(params)
    else:
        fg = FlatGradients(params, [shadow.source_of(p) for p in params] if shadow else None)
    if world > 1 and zs is None:
        if args.overlap:
            # bucketed all-reduce launched from inside backward on a side stream (captured in the step's CUDA graph).
            # Measured SLOWER at N = 2 (31.0 against 30.3 ms): the step's kernels are persistent, one CTA per SM with
            # ~220 KB of shared memory, so every SM an NCCL CTA sits on delays a whole CTA of theirs.  Off by default.
            fg.enable_overlap(world, n_buckets=args.buckets, stream=torch.cuda.Stream())
        elif args.allreduce_dtype == "bf16" and amp is not None:
            fg.enable_compressed_all_reduce(torch.bfloat16)
    net = model
    B = args.batch
    g = torch.Generator().manual_seed(rank)
    h_img = torch.randn(B, 3, IMG, IMG, generator=g).pin_memory()
    h_lab = (torch.rand(B, 1, IMG, IMG, generator=g) > 0.5).float().pin_memory()
    d_img, d_lab = h_img.to(dev), h_lab.to(dev)

    def fwd_bwd(img, lab):
        # autograd ASSIGNS fresh gradients (p.grad is None: no zero-fill + 679 accumulate kernels); they are
        # then packed into the flat buffer with one multi-tensor copy and the optimizer sees the flat views
        fg.begin()
        if zs:
            zs.begin()
        with torch.autocast("cuda", dtype=amp, enabled=amp is not None):
            outs = torch.func.functional_call(net, shadow.mapping(), (img,)) if shadow else net(img)
        loss = tsod_loss(outs, lab)
        loss.backward()
        fg.pack()
        if zs:
            zs.pack()
        return loss

    def reduce_grads():
        if zs:
            zs.reduce_scatter()
        fg.all_reduce(world)

    def opt_step():
        if zs:
            zs.update()                 # Adam on this rank's shard of the masters, rounded into its slice of the shadows
        opt.step()
        if shadow and not zs:
            shadow.refresh()            # one multi-tensor copy: the bf16 weights of the next forward

    def gather_weights():
        if zs:
            zs.all_gather()

    def step(img, lab):
        loss = fwd_bwd(img, lab)
        reduce_grads()
        opt_step()
        gather_weights()
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(args.warmup, 3)
    for _ in range(W):
        step(d_img, d_lab)
    sync()

    sampler = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- capture the training step in CUDA graphs: the step is static-shaped and launch-bound in eager
    # mode (~10^4 launches).  g_fb = zero grads + forward + loss + backward; g_opt = fused Adam step; the
    # gradient all-reduce (N > 1) runs between the two replays on the same stream.
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    g_fb = g_opt = None
    s_img, s_lab, s_loss = d_img.clone(), d_lab.clone(), None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step(s_img, s_lab)
            torch.cuda.current_stream().wait_stream(side)
            sync()
            g_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fb):
                s_loss = fwd_bwd(s_img, s_lab)
            g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_opt, pool=g_fb.pool()):
                opt_step()
            for _ in range(2):
                g_fb.replay(), reduce_grads(), g_opt.replay(), gather_weights()
            sync()
        except Exception as exc:
            import traceback
            traceback.print_exc()
            raise SystemExit(f"bench.py: CUDA graph capture failed ({type(exc).__name__}: {str(exc)[:300]}); "
                             f"re-run with --no-graph")
    graph = g_fb

    def graph_step():
        g_fb.replay()
        reduce_grads()
        g_opt.replay()
        gather_weights()

    # per-kernel device time INSIDE the graph-replayed step (CUPTI): what the roofline is computed from.
    # Single-process runs only: the profiled replays happen on one rank, so they must not contain a collective
    # (at N > 1 the step has the gradient all-reduce; the roofline then comes from the event-timed eager region).
    ktimes = None
    if graph is not None and world == 1:
        def local_step():
            g_fb.replay()
            g_opt.replay()
        try:
            ktimes = kernel_times(local_step, steps=3, path=args.kernel_table)
        except Exception as exc:  # no CUPTI: the roofline falls back to the event-timed eager region
            print(f"bench.py: torch.profiler unavailable ({type(exc).__name__}); roofline from CUDA events",
                  file=sys.stderr)

    # ---- region B (`value`): device-resident inputs
    sync()
    e0.record()
    for _ in range(args.steps):
        if graph is not None:
            graph_step()
        else:
            step(d_img, d_lab)
    e1.record()
    sync()
    ms_dev = e0.elapsed_time(e1)

    # ---- region C (`e2e`): host buffers in, loss out, every step
    sync()
    e0.record()
    for _ in range(args.steps):
        if graph is not None:
            s_img.copy_(h_img, non_blocking=True)
            s_lab.copy_(h_lab, non_blocking=True)
            graph_step()
            lv = s_loss.item()                          # device -> host read of the step's result
        else:
            img = h_img.to(dev, non_blocking=True)
            lab = h_lab.to(dev, non_blocking=True)
            lv = step(img, lab).item()
    e1.record()
    sync()
    ms_e2e = e0.elapsed_time(e1)
    # ---- region A (eager, instrumented; runs last so that nothing it creates is alive during capture): every launch of OUR kernels bracketed by CUDA events on the
    # launching stream -> per-kernel durations for the roofline; also the launch count per step
    _lib.prof = []
    n0 = _lib.launch_count()
    sync()
    e0.record()
    for it in range(args.steps):
        last = it == args.steps - 1
        if last:
            torch.cuda.profiler.start()     # cudaProfilerStart: `ncu --profile-from-start off` lists exactly
        loss = step(d_img, d_lab)           # this one eager step (a no-op without a profiler attached)
        if last:
            torch.cuda.profiler.stop()
    e1.record()
    sync()
    ms_eager = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    prof, _lib.prof = _lib.prof, None
    kern = summarize_prof(prof)
    loss_val = loss.item()

    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()

    # ---- after the timed regions: the replicas must still hold the same weights (a data-parallel step that lets them
    # drift would be measuring the wrong thing).  Checksums of the weights the next forward reads, compared over ranks.
    replica_check = None
    if world > 1:
        with torch.no_grad():
            ws = ([zs.w_flat] if zs else []) + [p for p in params]
            cs = torch.stack([torch.stack([w.double().sum(), w.double().abs().sum()]) for w in ws]).sum(0)
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        replica_check = {"identical": bool(torch.equal(lo, hi)), "weight_abs_sum": float(cs[1]),
                         "finite": bool(torch.isfinite(cs).all()), "loss": loss_val}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks \
            else (6650.0, "fallback (B200_PROFILING.md)")
        timed_in = "eager instrumented region (CUDA events around every launch of our kernels)"
        step_ms_for_share = ms_eager / args.steps
        if ktimes:
            # replace the event-timed durations (which include the CPU launch gap of eager mode) by the CUPTI
            # device times of the same launches inside the graph-replayed step; bytes/launch counts stay
            cupti = {}
            for name, (n, us) in ktimes.items():
                kind = kind_of(name)
                if kind is not None:
                    c = cupti.setdefault(kind, [0.0, 0.0])
                    c[0] += n
                    c[1] += us
            for kind, a in kern.items():
                c = cupti.get(kind)
                if c:
                    a["ms"] = c[1] * 1e-3 * args.steps
                    a["us_per_launch"] = c[1] / max(c[0], 1)
                    a["gbs"] = a["bytes"] / (a["ms"] * 1e-3) / 1e9
                    a["cupti"] = True
            timed_in = "graph-replayed step, per-launch device time from CUPTI activity records (torch.profiler)"
            step_ms_for_share = ms_dev / args.steps
        # the dominant kernel of the hot path proper (scan / selective scan / merge), by device time;
        # the ops either side of it (LayerNorm2d, depthwise conv, DCT) are listed under "kernels"
        path = [kk for kk in kern if kk.startswith(("ss2d_fused", "sscan", "cross_"))] or list(kern)
        dom = max(path, key=lambda kk: kern[kk]["ms"])
        traffic = None
        try:
            traffic = json.load(open(NCU_TRAFFIC_PATH)).get(dom)
        except (OSError, ValueError):
            pass
        k = kern[dom]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": k["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": k["gbs"] / peak, "traffic": traffic,
                    "traffic_source": "profiles/roofline_traffic.json (ncu --set full dram__bytes_read+write per "
                                      "launch, captured in an earlier run of the same kernels; not measured live)",
                    "peak_source": peak_src,
                    "launches": k["launches"], "us_per_launch": k["us_per_launch"],
                    "algorithmic_bytes_per_launch": k["bytes"] / k["launches"],
                    "share_of_step": (k["ms"] / args.steps) / step_ms_for_share,
                    "timed_in": timed_in if k.get("cupti") or not ktimes else
                    "eager instrumented region (CUDA events around every launch of our kernels)"}
        line = {
            "metric": METRIC, "value": world * B * args.steps / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": {"workload": "Tramba-V-TSOD training step 384x384, batch 4 per GPU (BASELINE config 3)",
                       "img_size": IMG, "global_batch": world * B, "parallelism": f"dp{world}",
                       "weights": "random-init, 111.44 M params", "optimizer": "Adam (fused)",
                       "collective": ("none (single GPU)" if world == 1 else
                                      (f"ZeRO-1: NCCL reduce-scatter (AVG) of the bf16 gradients ({zs.bytes_on_the_wire()[0]} "
                                       f"bytes in), Adam on 1/{world} of the fp32 masters per rank, all-gather of the bf16 "
                                       f"weights ({zs.bytes_on_the_wire()[1]} bytes out); fp32-only parameters: all-reduce of "
                                       f"{fg.flat.numel() * 4} bytes + replicated Adam") if zs else
                                      "NCCL all-reduce (AVG) of the flat gradient buffer, "
                                      + (f"{fg.wire.numel() * fg.wire.element_size()} bytes per step in bf16"
                                         if getattr(fg, "wire", None) is not None else f"{fg.flat.numel() * 4} bytes per step in fp32")
                                      + (f", {args.buckets} buckets launched from inside backward on a side stream"
                                         if args.overlap else ", one call between backward and Adam")),
                       "weight_precision": ("fp32 masters sharded over the ranks + bf16 shadows all-gathered after Adam" if zs else
                                            "fp32 masters + bf16 shadows refreshed after Adam") if shadow else "fp32, autocast casts per use",
                       "ss2d_path": "op-by-op kernels" if args.unfused else "fused core kernels",
                       "step_launch": ("CUDA graphs (fwd+bwd graph, reduce-scatter, Adam graph, all-gather)" if zs else "CUDA graphs (fwd+bwd graph, all-reduce, Adam graph)") if graph is not None else "eager",
                       "l2": "no explicit flush: a step streams several GB of activations (>> 126 MB L2)"},
            "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h_img.numel() * 4 + h_lab.numel() * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "replica_check": replica_check,
            "roofline": roofline,
            "kernels": {kk: {"launches": v["launches"], "us_per_launch": round(v["us_per_launch"], 2),
                             "gbs": round(v["gbs"], 1), "frac": round(v["gbs"] / peak, 4),
                             "share_of_step": round((v["ms"] / args.steps) / step_ms_for_share, 4)}
                        for kk, v in sorted(kern.items())},
            "kernels_timed_in": timed_in,
            "eager_ms_per_step": ms_eager / args.steps,
            "cuda_graph": graph is not None, "cuda_graph_error": graph_error,
            "loss": loss_val,
        }
        if world == 1 and not args.no_cpu_baseline:
            stepc, cores, sample = cpu_path_step_fn(1)
            stepc()                                     # warm-up: order tables, allocator
            t0 = time.perf_counter()
            stepc()
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": sample + "; 1 warm-up + 1 timed step"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------------------------
# BASELINE configs 2, 4, 5 (single GPU; the driver's headline is config 3 above): same JSON shape
# ----------------------------------------------------------------------------------------------
def _peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def _roofline_from(kern, ktimes, steps, step_ms, prefer):
    """kern: summarize_prof() of ONE instrumented eager pass (bytes / launch counts); ktimes: CUPTI device times of the
    graph-replayed pass (None -> the event times).  Returns (roofline dict, per-kernel dict)."""
    peak, peak_src = _peak()
    timed_in = "eager instrumented pass (CUDA events around every launch of our kernels)"
    if ktimes:
        cupti = {}
        for name, (n, us) in ktimes.items():
            kind = kind_of(name)
            if kind is not None:
                c = cupti.setdefault(kind, [0.0, 0.0])
                c[0] += n
                c[1] += us
        for kind, a in kern.items():
            c = cupti.get(kind)
            if c:
                a["ms"] = c[1] * 1e-3 * steps
                a["us_per_launch"] = c[1] / max(c[0], 1)
                a["gbs"] = a["bytes"] / (a["ms"] * 1e-3) / 1e9
        timed_in = "graph-replayed pass, per-launch device time from CUPTI activity records (torch.profiler)"
    path = [kk for kk in kern if kk.startswith(prefer)] or list(kern)
    dom = max(path, key=lambda kk: kern[kk]["ms"])
    k = kern[dom]
    roof = {"bound": "hbm", "kernel": dom, "achieved": k["gbs"], "peak": peak, "unit": "GB/s", "frac": k["gbs"] / peak,
            "traffic": None, "peak_source": peak_src, "launches": k["launches"], "us_per_launch": k["us_per_launch"],
            "algorithmic_bytes_per_launch": k["bytes"] / k["launches"],
            "share_of_step": (k["ms"] / steps) / step_ms, "timed_in": timed_in}
    kernels = {kk: {"launches": v["launches"], "us_per_launch": round(v["us_per_launch"], 2), "gbs": round(v["gbs"], 1),
                    "frac": round(v["gbs"] / peak, 4), "share_of_step": round((v["ms"] / steps) / step_ms, 4)}
               for kk, v in sorted(kern.items())}
    return roof, kernels


def run_forward_config(args, img, batch, workload, metric):
    """Configs 4 / 5: Tramba-V forward (eval, no_grad, bf16 autocast on bf16 shadow weights), CUDA-graph replayed.
    value: inputs resident; e2e: pinned host image batch copied in and the full-resolution mask read back each step."""
    import torch
    import tramba_b200  # noqa: F401
    from tramba_b200 import _lib
    from tramba_b200.dp import ShadowParams
    from tramba_b200.model import build_model
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = build_model(deep_supervision=False, img_size=img, dims=128).to(dev).eval()
    shadow = ShadowParams(model, torch.bfloat16)
    h_img = torch.randn(batch, 3, img, img).pin_memory()
    s_img = h_img.to(dev)

    def fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            out = torch.func.functional_call(model, shadow.mapping(), (s_img,))
        return out[-1] if isinstance(out, (list, tuple)) else out

    W = max(args.warmup, 3)
    for _ in range(W):
        out = fwd()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    t_wall0 = time.time()
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
    ktimes = None
    try:
        ktimes = kernel_times(g.replay, steps=3, path=args.kernel_table)
    except Exception as exc:  # noqa: BLE001
        print(f"bench.py: torch.profiler unavailable ({type(exc).__name__})", file=sys.stderr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms_dev = e0.elapsed_time(e1)
    h_out = torch.empty(out.shape, dtype=out.dtype).pin_memory()
    e0.record()
    for _ in range(args.steps):
        s_img.copy_(h_img, non_blocking=True)
        g.replay()
        h_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    _lib.prof = []
    n0 = _lib.launch_count()
    fwd()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - n0
    prof, _lib.prof = _lib.prof, None
    kern = summarize_prof(prof)
    clocks = sampler.stop(t_wall0, time.time())
    roof, kernels = _roofline_from(kern, ktimes, 1, ms_dev / args.steps, ("ss2d_fused", "sscan", "cross_"))
    line = {"metric": metric, "value": batch * args.steps / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": 1,
            "steps": args.steps, "warmup": W, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "img_size": img, "global_batch": batch, "parallelism": "dp1",
                       "weights": "random-init, bf16 shadows of fp32 masters", "step_launch": "CUDA graph",
                       "l2": "no explicit flush: one pass streams several GB of activations (>> 126 MB L2)"},
            "e2e": {"value": batch * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h_img.numel() * 4,
                    "d2h_bytes_per_step": h_out.numel() * h_out.element_size(), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches * args.steps, "clocks": clocks, "roofline": roof, "kernels": kernels,
            "finite": bool(torch.isfinite(out.float()).all()), "cpu_baseline": None}
    print(json.dumps(line), flush=True)
    return 0


def run_block_config(args):
    """Config 2: ONE SS2D block (d_model 128 -> d_inner 256, d_state 16, K = 4, batch 4, 96 x 96), forward + backward,
    fp32 like the reference trains (no AMP, train.py:29-99).  d_state 16 is outside the fused core (N = 1): the
    operator kernels run (cross-scan, selective scan fwd / bwd, cross-merge).  value = SS2D blocks fwd+bwd per second
    expressed as scan tokens/s; the roofline entry is the selective-scan backward, which at N = 16 is bound by the
    FP32 / MUFU issue rate rather than by HBM -- both fractions are what they are."""
    import torch
    import torch.nn as nn
    import tramba_b200
    from tramba_b200 import _lib, ops
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    B, Dm, H, N, K = 4, 128, 96, args.d_state, 4
    m = tramba_b200.SS2D(d_model=Dm, d_state=N, ssm_ratio=2.0, dt_rank="auto", act_layer=nn.SiLU, d_conv=3,
                         conv_bias=False, channel_first=True, disable_z=True, scan=ops.CrossScan, merge=ops.CrossMerge,
                         k_group=K).to(dev)
    h_x = torch.randn(B, Dm, H, H).pin_memory()
    x = h_x.to(dev).requires_grad_(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(xin):
        m.zero_grad(set_to_none=True)
        y = m(xin)
        y.sum().backward()
        return y

    W = max(args.warmup, 3)
    for _ in range(W):
        step(x)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    t_wall0 = time.time()
    tokens = B * K * 2 * Dm * H * H
    ts = []
    for _ in range(args.steps):
        flush.zero_()                                   # 256 MB > L2: every timed pass starts cold
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms_dev = sum(ts)
    te = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        xi = h_x.to(dev, non_blocking=True).requires_grad_(True)
        y = step(xi)
        _ = y.float().sum().item()                      # device -> host read of the step's result
        e1.record()
        torch.cuda.synchronize()
        te.append(e0.elapsed_time(e1))
    ms_e2e = sum(te)
    _lib.prof = []
    n0 = _lib.launch_count()
    step(x)
    torch.cuda.synchronize()
    launches = _lib.launch_count() - n0
    prof, _lib.prof = _lib.prof, None
    kern = summarize_prof(prof)
    clocks = sampler.stop(t_wall0, time.time())
    roof, kernels = _roofline_from(kern, None, 1, ms_dev / args.steps, ("sscan_bwd",))
    roof["note"] = ("N = 16: ~7 FP32 + 1 MUFU instructions per (token, state) exceed the 20 B/token HBM time (SURVEY 8d): "
                    "issue-bound by construction; the HBM fraction is reported for completeness")
    line = {"metric": "SS2D block fwd+bwd (d_state %d, K=4, batch 4, 96x96) scan tokens/sec" % N,
            "value": tokens * args.steps / (ms_dev * 1e-3), "unit": "tokens/s", "n_gpus": 1, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 2: single SS2D block fwd+bwd, batch 4, 96x96 tokens, d_state %d, K=4" % N,
                       "d_model": Dm, "d_inner": 2 * Dm, "global_batch": B, "parallelism": "dp1", "step_launch": "eager",
                       "l2": "flushed (256 MB write) before every timed step"},
            "e2e": {"value": tokens * args.steps / (ms_e2e * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": h_x.numel() * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches * args.steps, "clocks": clocks, "roofline": roof, "kernels": kernels,
            "cpu_baseline": None}
    print(json.dumps(line), flush=True)
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 2:
        return run_block_config(args)
    if args.config == 4:
        return run_forward_config(args, 384, 32, "Tramba-V-SOD inference 384x384, batch 32 per GPU (BASELINE config 4)",
                                  "Tramba-V-SOD 384x384 inference images/sec")
    if args.config == 5:
        return run_forward_config(args, 768, 8, "Tramba-V forward 768x768, batch 8: 192x192 maps, long helix / dilated "
                                  "scans, DCT2D at 192 (BASELINE config 5)", "Tramba-V 768x768 forward images/sec")
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
