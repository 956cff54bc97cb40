"""Data-parallel host logic of the training step: the path shards by image, so the only exchange
is the gradient all-reduce.  Every gradient lives in ONE flat fp32 buffer that is all-reduced once per step.

    fg = FlatGradients(model.parameters())
    fg.begin()                      # p.grad = None: autograd ASSIGNS fresh gradients (no zero-fill, no
    loss.backward()                 #                per-parameter accumulate kernels)
    fg.pack()                       # one multi-tensor copy into the flat buffer; p.grad = views of it
    fg.all_reduce(world)            # NCCL: AVG; gloo (CPU tests): SUM then scale
    (or fg.enable_overlap(world, n_buckets, side_stream) once: the all-reduce then runs bucket by bucket on a side
     stream DURING backward, and pack() only joins it)
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


def _grouped_copy(dsts, srcs):
    """torch._foreach_copy_ per (destination dtype, source dtype) group.  A list that mixes bf16 gradients (shadow
    weights) with fp32 ones (norms, A_logs, Ds, biases) misses the multi-tensor fast path and degrades to one copy
    kernel per tensor -- 673 launches per step for Tramba-V (measured: 314 cast kernels + 359 memcpys, 2.2 ms)."""
    groups = {}
    for d, s_ in zip(dsts, srcs):
        a, b = groups.setdefault((d.dtype, s_.dtype), ([], []))
        a.append(d)
        b.append(s_)
    for a, b in groups.values():
        torch._foreach_copy_(a, b)


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

    # ---- overlap of the gradient all-reduce with the backward pass (world > 1) -----------------------------------
    def enable_overlap(self, world, n_buckets=4, stream=None):
        """Splits the flat buffer into `n_buckets` contiguous ranges (by bytes, cut at parameter boundaries).  A
        post-accumulate hook on every gradient source counts arrivals; when the last gradient of a range has been
        produced, the range is packed (one multi-tensor copy) and all-reduced on a SIDE stream while autograd keeps
        running on the main one -- the decoder's gradients travel while the encoder's are still being computed.
        pack() then only joins the side stream.  `stream`: a torch.cuda.Stream, or None on CPU (gloo tests), where the
        same bucket logic runs inline.  Capturable in a CUDA graph (the side stream forks from / joins the capturing
        stream; NCCL collectives are graph-capturable)."""
        self.world = world
        self.side = stream
        total = self.flat.numel()
        target = (total + n_buckets - 1) // n_buckets
        self.buckets, cur, acc, off = [], [], 0, 0
        lo = 0
        for i, p in enumerate(self.params):
            cur.append(i)
            acc += p.numel()
            off += p.numel()
            if acc >= target or i == len(self.params) - 1:
                self.buckets.append({"idx": cur, "lo": lo, "hi": off})
                cur, acc, lo = [], 0, off
        self.bucket_of = {}
        for b, bk in enumerate(self.buckets):
            for i in bk["idx"]:
                self.bucket_of[i] = b
        self._arrived = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._hooks = []
        for i, src in enumerate(self.sources):
            self._hooks.append(src.register_post_accumulate_grad_hook(lambda t, i=i: self._on_grad(i)))
        self.overlap = True

    def _on_grad(self, i):
        if not getattr(self, "_armed", False):
            return
        b = self.bucket_of[i]
        self._arrived[b] += 1
        if self._arrived[b] == len(self.buckets[b]["idx"]) and not self._launched[b]:
            self._launch(b)

    def _launch(self, b):
        bk = self.buckets[b]
        self._launched[b] = True
        views = [self.views[i] for i in bk["idx"]]
        grads = [self.sources[i].grad for i in bk["idx"]]
        chunk = self.flat[bk["lo"]:bk["hi"]]

        def work():
            _grouped_copy(views, grads)
            self._reduce(chunk)
        if self.side is None:
            work()
        else:
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                work()

    def _reduce(self, t):
        if self.world <= 1:
            return
        if dist.get_backend() == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        else:                           # gloo has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.div_(self.world)

    def begin(self):
        for p, s in zip(self.params, self.sources):
            p.grad = None
            s.grad = None
        if getattr(self, "overlap", False):
            self._arrived = [0] * len(self.buckets)
            self._launched = [False] * len(self.buckets)
            self._armed = True

    def pack(self):
        if getattr(self, "overlap", False):
            # buckets whose gradients all arrived are already packed and reduced (or in flight on the side stream);
            # the rest (parameters that received no gradient this step) are completed here
            self._armed = False
            for b, bk in enumerate(self.buckets):
                if not self._launched[b]:
                    for i in bk["idx"]:
                        if self.sources[i].grad is None:
                            self.views[i].zero_()
                        else:
                            self.views[i].copy_(self.sources[i].grad)
                    self._launched[b] = True
                    chunk = self.flat[bk["lo"]:bk["hi"]]
                    if self.side is None:
                        self._reduce(chunk)
                    else:
                        self.side.wait_stream(torch.cuda.current_stream())
                        with torch.cuda.stream(self.side):
                            self._reduce(chunk)
            if self.side is not None:
                torch.cuda.current_stream().wait_stream(self.side)
            for v, p, s in zip(self.views, self.params, self.sources):
                if s is not p:
                    s.grad = None
                p.grad = v
            return
        dst = self.wire_views if getattr(self, "wire", None) is not None else self.views   # compressed: reduce first
        got = [(v, s.grad) for v, s in zip(dst, self.sources) if s.grad is not None]
        _grouped_copy([v for v, _ in got], [g for _, g in got])
        for v, w, p, s in zip(self.views, dst, self.params, self.sources):
            if s.grad is None:          # parameter unused in this step
                w.zero_()
            if s is not p:
                s.grad = None
            p.grad = v

    def enable_compressed_all_reduce(self, dtype=torch.bfloat16):
        """All-reduce the gradients in `dtype` (half the bytes on the wire): pack() writes them into a second flat
        buffer of that dtype, all_reduce() reduces it and widens it into the fp32 flat buffer the optimizer reads.  With
        bf16 shadow weights most gradients ARE bf16 already, so only the few fp32-owned parameters (norms, A_logs, Ds,
        biases) are rounded."""
        self.wire = torch.zeros_like(self.flat, dtype=dtype)
        self.wire_views, off = [], 0
        for p in self.params:
            self.wire_views.append(self.wire[off:off + p.numel()].view_as(p))
            off += p.numel()

    def all_reduce(self, world):
        if world <= 1 or getattr(self, "overlap", False):     # overlapped mode: done bucket by bucket during backward
            return
        if getattr(self, "wire", None) is not None:
            if dist.get_backend() == "nccl":
                dist.all_reduce(self.wire, op=dist.ReduceOp.AVG)
            else:
                dist.all_reduce(self.wire, op=dist.ReduceOp.SUM)
                self.wire.div_(world)
            self.flat.copy_(self.wire)
            return
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:                           # gloo has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(world)
