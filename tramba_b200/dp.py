"""Data-parallel host logic of the training step: the path shards by image, so the only exchange
is the gradient all-reduce.  Every gradient lives in ONE flat fp32 buffer that is all-reduced once per step.

    fg = FlatGradients(model.parameters())
    fg.begin()                      # p.grad = None: autograd ASSIGNS fresh gradients (no zero-fill, no
    loss.backward()                 #                per-parameter accumulate kernels)
    fg.pack()                       # one multi-tensor copy into the flat buffer; p.grad = views of it
    fg.all_reduce(world)            # NCCL: AVG; gloo (CPU tests): SUM then scale
    optimizer.step()                # sees the flat views

The reference has no distributed code (`run.py:46-50` only prints); this is build-side.
"""
import torch
import torch.distributed as dist


class FlatGradients:
    def __init__(self, params):
        self.params = [p for p in params]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=torch.float32)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            p.grad = self.views[-1]
            off += p.numel()

    def begin(self):
        for p in self.params:
            p.grad = None

    def pack(self):
        got = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None]
        if got:
            torch._foreach_copy_([v for v, _ in got], [g for _, g in got])
        for v, p in zip(self.views, self.params):
            if p.grad is None:          # parameter unused in this step
                v.zero_()
            p.grad = v

    def all_reduce(self, world):
        if world <= 1:
            return
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:                           # gloo has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(world)
