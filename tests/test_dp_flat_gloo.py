"""The data-parallel host logic bench.py uses (tramba_b200.dp.FlatGradients): world_size-2 gloo run on CPU.
After begin() -> backward -> pack() -> all_reduce(), every rank's flat buffer (and every p.grad view) holds the
MEAN of the per-rank gradients, and an optimizer step from the flat views keeps the replicas identical."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.GELU(), torch.nn.Conv2d(8, 4, 1),
                               torch.nn.Flatten(), torch.nn.Linear(4 * 8 * 8, 5), torch.nn.Linear(5, 1, bias=False))


def _worker(rank, world, port, out_dir, overlap=0):
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tramba_b200.dp import FlatGradients

    model = _model()
    params = list(model.parameters())
    opt = torch.optim.Adam(params, lr=1e-2)
    fg = FlatGradients(params)
    if overlap == -1:   # gradients reduced in bf16 on the wire, widened into the fp32 buffer the optimizer reads
        fg.enable_compressed_all_reduce(torch.bfloat16)
        overlap = 0
        wire = True
    else:
        wire = False
    if overlap:      # bucketed all-reduce launched from the gradient hooks DURING backward (inline on CPU: stream=None)
        fg.enable_overlap(world, n_buckets=overlap, stream=None)
        assert len(fg.buckets) >= 2 and sum(len(b["idx"]) for b in fg.buckets) == len(params)
        assert fg.buckets[0]["lo"] == 0 and fg.buckets[-1]["hi"] == fg.flat.numel()
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, fg.views))
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(2, 3, 8, 8, generator=g)
    worst = 0.0
    for _ in range(2):                                        # two steps: the views must survive a step
        fg.begin()
        assert all(p.grad is None for p in params)
        if overlap:
            local = []
            hooks = [p.register_post_accumulate_grad_hook(lambda t, local=local: local.append((id(t), t.grad.clone())))
                     for p in params]      # registered AFTER fg's hooks?  order is irrelevant: grads are not mutated
        model(x).square().mean().backward()
        if overlap:
            for h in hooks:
                h.remove()
            byid = dict(local)
            local = [byid[id(p)] for p in params]
            assert all(fg._launched[:-1]) or all(fg._launched)   # every complete bucket left during backward
        else:
            local = [p.grad.clone() for p in params]
        fg.pack()
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, fg.views))
        if not overlap and not wire:
            assert all(torch.equal(p.grad, l) for p, l in zip(params, local))
        fg.all_reduce(world)
        for p, l in zip(params, local):
            mean = l.bfloat16().float() if wire else l.clone()     # bf16 wire: every rank's gradient is rounded once
            dist.all_reduce(mean)
            mean /= world
            err = float((mean - p.grad).abs().max())
            worst = max(worst, err / (float(mean.abs().max()) + 1e-12) if wire else err)   # bf16 wire: relative error
        opt.step()
    w = torch.cat([p.detach().flatten() for p in params])
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    same = all(torch.equal(ws[0], t) for t in ws)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([worst, float(same)]))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_flat_gradients_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        worst, same = np.load(tmp_path / f"r{r}.npy")
        assert worst < 1e-6, worst
        assert same == 1.0


@pytest.mark.timeout(300)
def test_flat_gradients_overlapped_buckets_world2_gloo(tmp_path):
    """The overlapped mode (bench.py at N > 1): gradient hooks pack + all-reduce each bucket as soon as its last
    gradient exists; pack() completes whatever is left.  Same result as the blocking path: the mean on every rank."""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), 3), nprocs=world, join=True)
    for r in range(world):
        worst, same = np.load(tmp_path / f"r{r}.npy")
        assert worst < 1e-6, worst
        assert same == 1.0


@pytest.mark.timeout(300)
def test_flat_gradients_bf16_wire_world2_gloo(tmp_path):
    """bench.py's default at N > 1 in the bf16 step: the all-reduce runs on a bf16 copy of the gradients (half the
    bytes), the optimizer reads the widened fp32 buffer.  Mean of the per-rank (bf16-rounded) gradients within bf16
    resolution on every rank; replicas stay identical."""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), -1), nprocs=world, join=True)
    for r in range(world):
        worst, same = np.load(tmp_path / f"r{r}.npy")
        assert worst < 2.0 ** -6, worst       # a few bf16 ulps of the tensor's largest entry
        assert same == 1.0


def _zero_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.GroupNorm(2, 8), torch.nn.GELU(),
                               torch.nn.Conv2d(8, 4, 1), torch.nn.Flatten(), torch.nn.Linear(4 * 8 * 8, 5),
                               torch.nn.Linear(5, 1, bias=False))


def _zero_worker(rank, world, port, out_dir, wire):
    """ShardedStep (reduce-scatter -> Adam on the rank's shard of the fp32 masters -> all-gather of the bf16 shadows)
    against the replicated step (all-reduce of the flat fp32 gradients -> Adam on every parameter -> refresh)."""
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tramba_b200.dp import FlatGradients, ShadowParams, ShardedStep

    def groups_of(model):   # two learning rates, like the encoder / decoder split of train.py:266-280
        ps = list(model.parameters())
        return [(ps[:4], 1e-3), (ps[4:], 1e-2)]

    # ---- replicated reference
    ref = _zero_model()
    ref_sh = ShadowParams(ref, torch.bfloat16)
    ref_params = list(ref.parameters())
    ref_fg = FlatGradients(ref_params, [ref_sh.source_of(p) for p in ref_params])
    ref_opt = torch.optim.Adam([{"params": ps, "lr": lr} for ps, lr in groups_of(ref)])
    # ---- sharded
    net = _zero_model()
    sh = ShadowParams(net, torch.bfloat16)
    groups = groups_of(net)
    zs = ShardedStep(sh, groups, world, rank, wire_dtype=wire)
    assert zs.n >= sum(m.numel() for m in sh.masters) and zs.n_pad % (8 * world) == 0
    assert all((s.data_ptr() - zs.w_flat.data_ptr()) % 256 == 0 for s in sh.shadows)   # what cuDNN / cuBLAS read
    assert all(s.data_ptr() >= zs.w_flat.data_ptr() for s in sh.shadows)          # shadows are views of the flat buffer
    has = {id(m) for m in sh.masters}
    rest = [(p, lr) for ps, lr in groups for p in ps if id(p) not in has]       # fp32-only parameters: replicated
    assert len(rest) == 2                                                        # the GroupNorm's weight and bias
    fg = FlatGradients([p for p, _ in rest])
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in rest])

    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(2, 3, 8, 8, generator=g)
    worst = 0.0
    for _ in range(3):
        ref_fg.begin()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            torch.func.functional_call(ref, ref_sh.mapping(), (x,)).float().square().mean().backward()
        ref_fg.pack()
        ref_fg.all_reduce(world)
        ref_opt.step()
        ref_sh.refresh()

        fg.begin()
        zs.begin()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            torch.func.functional_call(net, sh.mapping(), (x,)).float().square().mean().backward()
        fg.pack()
        zs.pack()
        zs.reduce_scatter()
        fg.all_reduce(world)
        zs.update()
        opt.step()
        zs.all_gather()
        for a, b in zip(sh.shadows, ref_sh.shadows):
            worst = max(worst, float((a.float() - b.float()).abs().max()) / (float(b.float().abs().max()) + 1e-12))
    zs.gather_masters()
    exact = all(torch.equal(a, b) for a, b in zip(net.parameters(), ref.parameters()))
    close = max(float((a - b).abs().max()) for a, b in zip(net.parameters(), ref.parameters()))
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([worst, float(exact), close]))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("wire", ["f32", "bf16"])
def test_sharded_step_matches_replicated_step_world2_gloo(tmp_path, wire):
    """ZeRO-1 step of bench.py at N > 1.  fp32 on the wire: bit-identical weights (shadows AND gathered fp32 masters)
    to the replicated all-reduce + full Adam step after 3 steps, with the shard boundary and the learning-rate
    boundary inside parameters.  bf16 on the wire (bench.py's default: the gradients of shadowed weights ARE bf16):
    the averaged gradient is rounded once more, so the weights agree to bf16 resolution."""
    world = 2
    mp.spawn(_zero_worker, args=(world, _free_port(), str(tmp_path), torch.float32 if wire == "f32" else torch.bfloat16),
             nprocs=world, join=True)
    for r in range(world):
        worst, exact, close = np.load(tmp_path / f"r{r}.npy")
        if wire == "f32":
            assert exact == 1.0 and worst == 0.0, (worst, exact, close)
        else:
            assert worst < 2.0 ** -6 and close < 1e-2, (worst, close)


def test_flat_gradients_unused_parameter_is_zeroed():
    sys.path.insert(0, os.path.dirname(HERE))
    from tramba_b200.dp import FlatGradients
    a, b = torch.nn.Parameter(torch.ones(3)), torch.nn.Parameter(torch.ones(2))
    fg = FlatGradients([a, b])
    fg.flat.fill_(7.0)
    fg.begin()
    (a * 2).sum().backward()                                  # b unused
    fg.pack()
    assert torch.equal(a.grad, torch.full((3,), 2.0)) and torch.equal(b.grad, torch.zeros(2))
    assert torch.equal(fg.flat, torch.tensor([2.0, 2.0, 2.0, 0.0, 0.0]))


def test_shadow_params_reproduce_autocast_gradients_and_updates():
    """ShadowParams + FlatGradients (what bench.py's bf16 step uses) against plain autocast on the same model: the bf16
    shadows are the values autocast would cast to, so outputs, gradients (after widening into the flat fp32 buffer)
    and the weights after two Adam steps are IDENTICAL bit for bit; parameters consumed in fp32 keep no shadow."""
    sys.path.insert(0, os.path.dirname(HERE))
    from tramba_b200.dp import FlatGradients, ShadowParams

    def make():
        torch.manual_seed(0)
        return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.GroupNorm(2, 8), torch.nn.GELU(),
                                   torch.nn.Conv2d(8, 8, 3, padding=1, groups=8),      # depthwise: stays fp32-owned
                                   torch.nn.Conv2d(8, 4, 1), torch.nn.Flatten(), torch.nn.Linear(4 * 8 * 8, 5))
    x = torch.randn(2, 3, 8, 8, generator=torch.Generator().manual_seed(3))
    ref, new = make(), make()
    opt_r = torch.optim.Adam(ref.parameters(), lr=1e-2)
    params = list(new.parameters())
    shadow = ShadowParams(new, torch.bfloat16)
    assert len(shadow.shadows) == 6                                   # conv0 w,b; conv 1x1 w,b; linear w,b
    assert all(s.dtype == torch.bfloat16 for s in shadow.shadows)
    assert shadow.source_of(new[3].weight) is new[3].weight           # depthwise conv: no shadow
    fg = FlatGradients(params, [shadow.source_of(p) for p in params])
    opt_n = torch.optim.Adam(params, lr=1e-2)
    for _ in range(2):
        opt_r.zero_grad(set_to_none=True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            yr = ref(x)
        yr.float().square().mean().backward()
        opt_r.step()
        fg.begin()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            yn = torch.func.functional_call(new, shadow.mapping(), (x,))
        yn.float().square().mean().backward()
        fg.pack()
        assert torch.equal(yr, yn)
        assert all(p.grad.dtype == torch.float32 and p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, fg.views))
        opt_n.step()
        shadow.refresh()
    for a, b in zip(ref.parameters(), new.parameters()):
        assert torch.equal(a, b)
    for m, s in zip(shadow.masters, shadow.shadows):
        assert torch.equal(m.detach().to(torch.bfloat16), s.detach())
