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

    @torch.no_grad()
    def rebase(self, flat, order, offsets):
        """Re-creates the shadows as views of `flat` (shadow order[j] at element offsets[j]): what ShardedStep
        all-gathers into."""
        for i, off in zip(order, offsets):
            n = self.masters[i].numel()
            v = flat[off:off + n].view_as(self.masters[i])
            v.copy_(self.shadows[i])
            self.shadows[i] = v.detach().requires_grad_(True)
        self._map = dict(zip(self.names, self.shadows))
        self._of = {id(p): s for p, s in zip(self.masters, self.shadows)}


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


class ShardedStep:
    """ZeRO-1 over the data-parallel ranks for the parameters that have a low-precision shadow (convolution, linear
    and projection weights: > 99 % of Tramba-V's 111 M parameters):

        backward                     leaves the (bf16) gradients on the shadows
        pack()                       one multi-tensor copy into a flat bf16 gradient buffer
        reduce_scatter()             NCCL reduce-scatter (AVG): every rank receives 1/N of the averaged gradients
        update()                     Adam on THIS RANK'S 1/N shard of the fp32 masters (+ its Adam moments), then the
                                     shard is rounded into its slice of the flat bf16 shadow buffer
        all_gather()                 NCCL all-gather of the bf16 shadows: the weights of the next forward

    Against "all-reduce the fp32 gradients, every rank runs the whole Adam step": the same bytes as a bf16 all-reduce
    on the wire (half those of the fp32 one), and the optimizer's HBM traffic (16 bytes per parameter read + 12
    written) drops to 1/N per rank -- at N = 8 that saves more time than the collectives cost.  The fp32 masters and
    moments exist only as shards (gather_masters() rebuilds the full fp32 parameters, e.g. for a checkpoint).  The few
    fp32-only parameters (norms, depthwise convs, A_logs, Ds, dt biases) stay replicated: FlatGradients + all-reduce +
    the caller's optimizer, as at N = 1.

    groups: list of (parameters, lr) in the order the caller wants them laid out (Tramba-V: encoder 1e-5, decoder
    1e-4, train.py:266-280); parameters without a shadow are ignored here.  The shadows become views of ONE flat
    buffer (shadow.rebase), so that the all-gather lands directly in the tensors the model is evaluated with.
    """

    ALIGN = 128      # elements between parameter starts in the flat buffers (256 bytes in bf16)

    def __init__(self, shadow, groups, world, rank, wire_dtype=None, betas=(0.9, 0.999), eps=1e-8, fused=None,
                 capturable=False):
        self.world, self.rank, self.shadow = world, rank, shadow
        dev = shadow.shadows[0].device
        sdt = shadow.shadows[0].dtype
        self.wire_dtype = wire_dtype or sdt
        has = {id(m): i for i, m in enumerate(shadow.masters)}
        # every parameter starts 256-byte aligned in the flat bf16 buffer: the shadows are what cuDNN / cuBLAS read, and
        # a weight pointer that is not 16-byte aligned sends cuDNN to slower kernels (measured: + 0.9 ms per step)
        A = self.ALIGN
        self.order, self.offs, self.bounds, n = [], [], [], 0   # shadow indices in layout order, their offsets; lr groups
        for params, lr in groups:
            lo = n
            for p in params:
                if id(p) in has:
                    self.order.append(has[id(p)])
                    self.offs.append(n)
                    n += (p.numel() + A - 1) // A * A
            self.bounds.append((lo, n, lr))
        assert len(self.order) == len(shadow.masters), "every shadowed parameter must be in exactly one group"
        self.n = n
        q = world * 8                                  # shards start 16-byte aligned in bf16 and fp32
        self.n_pad = (n + q - 1) // q * q
        self.ns = self.n_pad // world                  # elements per shard
        lo = rank * self.ns
        self.w_flat = torch.zeros(self.n_pad, device=dev, dtype=sdt)              # bf16 weights (shadows = views)
        self.g_flat = torch.zeros(self.n_pad, device=dev, dtype=self.wire_dtype)  # gradients on the wire
        self.g_shard = torch.zeros(self.ns, device=dev, dtype=self.wire_dtype)    # averaged gradients of my shard
        self.m_shard = torch.zeros(self.ns, device=dev, dtype=torch.float32)      # fp32 masters of my shard
        self.m_grad = torch.zeros(self.ns, device=dev, dtype=torch.float32)
        self.g_views = []
        masters_flat = torch.zeros(self.n_pad, device=dev, dtype=torch.float32)
        for i, off in zip(self.order, self.offs):
            m = shadow.masters[i]
            masters_flat[off:off + m.numel()].copy_(m.detach().reshape(-1))
            self.g_views.append(self.g_flat[off:off + m.numel()].view_as(m))
        self.m_shard.copy_(masters_flat[lo:lo + self.ns])
        shadow.rebase(self.w_flat, self.order, self.offs)
        # this rank's shard, cut at the learning-rate boundaries: one optimizer "parameter" per piece
        pgroups = []
        for glo, ghi, lr in self.bounds:
            a, b = max(glo, lo), min(ghi, lo + self.ns)
            if b > a:
                piece = self.m_shard[a - lo:b - lo]
                piece.grad = self.m_grad[a - lo:b - lo]
                pgroups.append({"params": [piece], "lr": lr})
        kw = {} if fused is None else {"fused": fused}
        self.opt = torch.optim.Adam(pgroups, betas=betas, eps=eps, capturable=capturable, **kw) if pgroups else None

    def bytes_on_the_wire(self):
        """(reduce-scatter input bytes, all-gather output bytes) per rank and step."""
        return self.g_flat.numel() * self.g_flat.element_size(), self.w_flat.numel() * self.w_flat.element_size()

    def begin(self):
        for s in self.shadow.shadows:
            s.grad = None

    def pack(self):
        srcs = [self.shadow.shadows[i] for i in self.order]
        got = [(v, s.grad) for v, s in zip(self.g_views, srcs) if s.grad is not None]
        _grouped_copy([v for v, _ in got], [g for _, g in got])
        for v, s in zip(self.g_views, srcs):
            if s.grad is None:
                v.zero_()
            s.grad = None

    def reduce_scatter(self):
        if self.world == 1:
            self.g_shard.copy_(self.g_flat)
        elif dist.get_backend() == "nccl":
            dist.reduce_scatter_tensor(self.g_shard, self.g_flat, op=dist.ReduceOp.AVG)
        else:                                          # gloo (CPU tests): no reduce-scatter, no AVG
            dist.all_reduce(self.g_flat, op=dist.ReduceOp.SUM)
            lo = self.rank * self.ns
            self.g_shard.copy_(self.g_flat[lo:lo + self.ns]).div_(self.world)

    @torch.no_grad()
    def update(self):
        self.m_grad.copy_(self.g_shard)                # widen: the optimizer reads fp32
        if self.opt is not None:
            self.opt.step()
        lo = self.rank * self.ns
        self.w_flat[lo:lo + self.ns].copy_(self.m_shard)    # round my shard into the flat bf16 weights

    def all_gather(self):
        if self.world > 1:
            lo = self.rank * self.ns
            dist.all_gather_into_tensor(self.w_flat, self.w_flat[lo:lo + self.ns])

    @torch.no_grad()
    def gather_masters(self):
        """Rebuilds the full fp32 master parameters (checkpoint surface) from the shards."""
        full = torch.empty(self.n_pad, device=self.m_shard.device, dtype=torch.float32)
        if self.world > 1:
            dist.all_gather_into_tensor(full, self.m_shard)
        else:
            full.copy_(self.m_shard)
        for i, off in zip(self.order, self.offs):
            m = self.shadow.masters[i]
            m.copy_(full[off:off + m.numel()].view_as(m))
