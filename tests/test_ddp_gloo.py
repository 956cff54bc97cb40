"""N > 1 host logic on CPU: world_size-2 gloo DDP over the Tramba-V harness (oracle operators
injected, tiny width).  The hot path shards by image with no data-path collective, so
what must hold is: (i) the custom autograd operators work under DistributedDataParallel
(static graph, all 673 parameters receive gradients), (ii) after backward every rank holds the
mean of the per-rank gradients, (iii) the max-over-ranks timing reduction bench.py uses."""
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


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _det_weights import det_fill, det_input
    from _oracle_ops import oracleize
    from tramba_b200.model import build_model, tsod_loss

    model = oracleize(det_fill(build_model(img_size=384, dims=8)), f32=True).eval()  # eval: DropPath off
    x = det_input((1, 3, 384, 384), 100 + rank)
    y = (det_input((1, 1, 384, 384), 200 + rank) > 0).float()
    # single-process gradient of this rank's shard
    tsod_loss(model(x), y).backward()
    local = [p.grad.clone() for p in model.parameters()]
    assert all(g is not None for g in local)
    model.zero_grad()
    ddp = torch.nn.parallel.DistributedDataParallel(model, static_graph=True)
    tsod_loss(ddp(x), y).backward()
    synced = [p.grad.clone() for p in model.parameters()]
    # reference reduction: mean over ranks of the local gradients
    worst = 0.0
    for g_local, g_sync in zip(local, synced):
        mean = g_local.clone()
        dist.all_reduce(mean)
        mean /= world
        worst = max(worst, float((mean - g_sync).abs().max() / (mean.abs().max() + 1e-12)))
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([worst, float(t)]))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_ddp_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        worst, tmax = np.load(tmp_path / f"r{r}.npy")
        assert worst < 1e-5, worst
        assert tmax == 2.0
