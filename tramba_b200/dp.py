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


class ShadowParams:
    """Low-precision SHADOW copies of the parameters that autocast would otherwise re-cast on every use.

    Under `torch.autocast(bf16)` every convolution / linear casts its fp32 weight (and bias) to bf16 in the forward
    and casts the bf16 weight gradient back in the backward: ~700 tiny kernels per Tramba-V step (7 % of it).  The
    same values result from keeping one bf16 copy per such parameter, refreshed by ONE multi-tensor copy after the
    optimizer step; the model is evaluated with `torch.func.functional_call(model, shadow.mapping(), ...)` and
    autograd leaves the (bf16) gradients on the shadows, from where FlatGradients widens them into the fp32 flat
    buffer (one multi-tensor copy).  Parameters consumed in fp32 (LayerNorm, depthwise convs on the library's own
    kernels, A_logs / Ds / dt_projs_bias of SS2D) have no shadow."""

    def __init__(self, model, dtype=torch.bfloat16):
        import torch.nn as nn
        from .ss2d import SS2D
        picked = {}
        for mname, m in model.named_modules():
            local = []
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
                depthwise = isinstance(m, nn.Conv2d) and m.groups > 1 and m.groups == m.in_channels
                if not depthwise:
                    local = [n for n, _ in m.named_parameters(recurse=False)]
            elif isinstance(m, SS2D):
                local = ["x_proj_weight", "dt_projs_weight"]       # operands of the projection GEMMs
            for n in local:
                picked[f"{mname}.{n}" if mname else n] = getattr(m, n)
        self.names = list(picked)
        self.masters = [picked[n] for n in self.names]
        self.shadows = [p.detach().to(dtype).requires_grad_(True) for p in self.masters]
        self._map = dict(zip(self.names, self.shadows))
        self._of = {id(p): s for p, s in zip(self.masters, self.shadows)}

    def mapping(self):
        return self._map

    def source_of(self, p):
        """The tensor whose .grad autograd fills for parameter p: its shadow if it has one."""
        return self._of.get(id(p), p)

    @torch.no_grad()
    def refresh(self):
        if self.shadows:
            torch._foreach_copy_(self.shadows, self.masters)


class FlatGradients:
    def __init__(self, params, sources=None):
        """sources[i]: the tensor on which autograd leaves parameter i's gradient (the parameter itself, or its
        low-precision shadow -- ShadowParams.source_of)."""
        self.params = [p for p in params]
        self.sources = list(sources) if sources is not None else self.params
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=torch.float32)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            p.grad = self.views[-1]
            off += p.numel()

    def begin(self):
        for p, s in zip(self.params, self.sources):
            p.grad = None
            s.grad = None

    def pack(self):
        got = [(v, s.grad) for v, s in zip(self.views, self.sources) if s.grad is not None]
        if got:
            torch._foreach_copy_([v for v, _ in got], [g for _, g in got])
        for v, p, s in zip(self.views, self.params, self.sources):
            if s.grad is None:          # parameter unused in this step
                v.zero_()
            if s is not p:
                s.grad = None
            p.grad = v

    def all_reduce(self, world):
        if world <= 1:
            return
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:                           # gloo has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(world)
