"""Import shim that makes /root/reference importable on a CPU-only box.

Used ONLY by tests/golden/make_golden.py (fixture generation, build container) and by the
optional reference cross-check tests that skip when /root/reference is absent.  Nothing here
copies reference source; it only patches the environment the reference expects:
  * torch.Tensor.cuda / nn.Module.cuda -> identity (module-scope .cuda() at csms6s.py:374-387)
  * a 3-symbol fake `timm.models.layers` (DropPath, trunc_normal_, to_2tuple)
"""
import os
import sys
import types

REF = os.environ.get("TRAMBA_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "Models", "SS2D"))


def install():
    import torch
    import torch.nn as nn

    if getattr(install, "_done", False):
        return
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob = drop_prob
            self.scale_by_keep = scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            mask = x.new_empty(shape).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                mask.div_(keep)
            return x * mask

    def to_2tuple(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = DropPath
    layers.trunc_normal_ = nn.init.trunc_normal_
    layers.to_2tuple = to_2tuple
    timm.models = models
    models.layers = layers
    sys.modules.setdefault("timm", timm)
    sys.modules.setdefault("timm.models", models)
    sys.modules.setdefault("timm.models.layers", layers)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    install._done = True
