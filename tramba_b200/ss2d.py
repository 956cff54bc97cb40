"""SS2D module -- same constructor, parameter names/shapes and forward contract as the
reference (Models/vmamba.py:18-323), running its core on the sm_100a operators.

State-dict keys: in_proj.weight, conv2d.weight, x_proj_weight (K, R+2N, D), dt_projs_weight
(K, D, R), dt_projs_bias (K, D), A_logs (K*D, N), Ds (K*D), out_norm.{weight,bias},
out_proj.weight -- reference checkpoints load unchanged.

The in/out projections, depthwise conv, x_proj/dt_proj and the LayerNorm stay PyTorch ops, as
in the reference (north_star); cross-scan, selective scan and cross-merge are ours.
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dwconv, fused, norm, ops


_PROJ_KERNEL = os.environ.get("TRAMBA_PROJ_GATHER", "kernel") != "torch"  # tramba_proj_gather vs torch.gather + split
_PROJ_GEMM = os.environ.get("TRAMBA_PROJ", "gemm") == "gemm"   # x_proj / dt_proj as batched GEMMs vs 1x1 conv1d


def _pointwise(x, weight, bias=None):
    """1x1 convolution of an NCHW map as ONE batched GEMM on the (B, C, L) view: the same function as
    F.conv2d(x, weight[:, :, None, None], bias) without cuDNN's NCHW<->NHWC round trips."""
    B, C, H, W = x.shape
    y = torch.bmm(weight.unsqueeze(0).expand(B, -1, -1), x.reshape(B, C, H * W))
    if bias is not None:
        y = y + bias.view(1, -1, 1)
    return y.view(B, -1, H, W)


class Linear2d(nn.Linear):
    """1x1 convolution with nn.Linear parameters (Models/modules.py:10-19)."""

    as_gemm = os.environ.get("TRAMBA_LINEAR2D", "conv") == "gemm"   # measured: cuDNN's 1x1 conv is faster (66.8 vs 72.2 ms/step)

    def forward(self, x):
        if self.as_gemm:
            return _pointwise(x, self.weight, self.bias)
        return F.conv2d(x, self.weight[:, :, None, None], self.bias)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        key = prefix + "weight"
        if key in state_dict:
            state_dict[key] = state_dict[key].view(self.weight.shape)
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class LayerNorm2d(nn.LayerNorm):
    """LayerNorm over the channel dim of an NCHW tensor (Models/modules.py:22-27).

    On a CUDA device this is ONE kernel (tramba_b200.norm.layer_norm2d) instead of permute ->
    F.layer_norm -> permute.  `act="gelu"` fuses the activation that follows SS2D.out_norm
    (Models/vmamba.py:287).  Output dtype = what the reference's ops produce: fp32 under autocast
    (F.layer_norm is an autocast-to-fp32 op) unless `to_dtype` asks for the consumer's dtype."""

    fused = os.environ.get("TRAMBA_LAYERNORM2D", "fused") != "torch"

    def forward(self, x, act="none", to_dtype=None):
        if self.fused and norm.supported(x, to_dtype):
            out = to_dtype or (torch.float32 if torch.is_autocast_enabled() else x.dtype)
            return norm.layer_norm2d(x, self.weight, self.bias, self.eps, act, out)
        x = x.permute(0, 2, 3, 1).contiguous()
        x = F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        x = x.permute(0, 3, 1, 2).contiguous()
        if to_dtype is not None:
            x = x.to(to_dtype)
        return F.gelu(x) if act == "gelu" else x


def _dt_init(dt_rank, d_inner, dt_scale=1.0, dt_init="random", dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4):
    """Models/mamba_init.py:7-31."""
    proj = nn.Linear(dt_rank, d_inner, bias=True)
    std = dt_rank ** -0.5 * dt_scale
    if dt_init == "constant":
        nn.init.constant_(proj.weight, std)
    elif dt_init == "random":
        nn.init.uniform_(proj.weight, -std, std)
    else:
        raise NotImplementedError
    dt = torch.exp(torch.rand(d_inner) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min)).clamp(
        min=dt_init_floor)
    with torch.no_grad():
        proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))  # softplus^-1
    return proj


def _a_log_init(d_state, d_inner, copies):
    """Models/mamba_init.py:34-48: S4D-real, A = 1..N per row, K copies merged."""
    a = torch.arange(1, d_state + 1, dtype=torch.float32).repeat(copies * d_inner, 1).contiguous()
    p = nn.Parameter(torch.log(a))
    p._no_weight_decay = True
    return p


def _d_init(d_inner, copies):
    """Models/mamba_init.py:51-60."""
    p = nn.Parameter(torch.ones(copies * d_inner))
    p._no_weight_decay = True
    return p


def _adopt(cls):
    """A scan / merge class that is not ours but carries the name of one of the reference's operators
    (csms6s.py:13-737) -> our class of that name; anything else (ours, custom pairs, unknown names) unchanged."""
    name = getattr(cls, "__name__", "")
    if getattr(cls, "__module__", "").startswith("tramba_b200") or not name.startswith(("CrossScan", "CrossMerge")):
        return cls
    ours = getattr(ops, name, None)
    return ours if isinstance(ours, type) and issubclass(ours, torch.autograd.Function) else cls


class SS2D(nn.Module):
    def __init__(
            self,
            d_model=96, d_state=16, ssm_ratio=2.0, dt_rank="auto", act_layer=nn.SiLU,
            d_conv=3, conv_bias=False,
            dropout=0.0, bias=False,
            dt_min=0.001, dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, initialize="v0",
            forward_type="v2", channel_first=False,
            disable_z=True,
            scan=ops.CrossScan,
            merge=ops.CrossMerge,
            k_group=4,
            **kwargs,
    ):
        super().__init__()
        d_inner = int(ssm_ratio * d_model)
        dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.channel_first = channel_first
        self.with_dconv = d_conv > 1
        self.disable_z = disable_z
        self.k_group = k_group
        # The reference's module tree passes ITS classes (`scan=CrossScan_Line`, Models/vmamba.py:666-667,
        # freq_mamba.py:22-33).  Same-named classes are the same operators: adopt ours, so that a module-level
        # swap reaches the kernels (and the fused core) without also patching the operator names.
        self.scan, self.merge = _adopt(scan), _adopt(merge)
        # bound like the reference's partial(forward_corev2, SelectiveScan=SelectiveScanOflex, ...)
        # (Models/vmamba.py:65-66); tests swap in the CPU oracle operator here
        self.selective_scan = ops.SelectiveScanOflex
        # fused=True: run scan -> selective scan -> merge as one kernel on image-ordered operands when the
        # (scan, merge) pair is one of ours and the shape is covered; False: the reference's op-by-op path
        self.fused = kwargs.get("fused", True)
        self.fused_dwconv = kwargs.get("fused_dwconv", os.environ.get("TRAMBA_DWCONV", "fused") != "torch")
        Linear = Linear2d if channel_first else nn.Linear
        LayerNorm = LayerNorm2d if channel_first else nn.LayerNorm

        self.out_norm = LayerNorm(d_inner)
        self.in_proj = Linear(d_model, d_inner if disable_z else d_inner * 2, bias=bias)
        self.act = act_layer()
        if self.with_dconv:
            self.conv2d = nn.Conv2d(d_inner, d_inner, groups=d_inner, bias=conv_bias, kernel_size=d_conv,
                                    padding=(d_conv - 1) // 2)
        x_proj = [nn.Linear(d_inner, dt_rank + d_state * 2, bias=False) for _ in range(k_group)]
        self.x_proj_weight = nn.Parameter(torch.stack([t.weight for t in x_proj], dim=0))  # (K, R+2N, D)
        self.out_act = nn.GELU()
        self.out_proj = Linear(d_inner, d_model, bias=bias)
        self.dropout = nn.Dropout(dropout) if dropout > 0.0 else nn.Identity()
        if initialize in ["v0"]:
            dt_projs = [_dt_init(dt_rank, d_inner, dt_scale, dt_init, dt_min, dt_max, dt_init_floor)
                        for _ in range(k_group)]
            self.dt_projs_weight = nn.Parameter(torch.stack([t.weight for t in dt_projs], dim=0))  # (K, D, R)
            self.dt_projs_bias = nn.Parameter(torch.stack([t.bias for t in dt_projs], dim=0))      # (K, D)
            self.A_logs = _a_log_init(d_state, d_inner, copies=k_group)                            # (K*D, N)
            self.Ds = _d_init(d_inner, copies=k_group)                                             # (K*D)

    def _fused_order(self, x):
        """(order, param) when this call can take the fused kernel, else None."""
        order = getattr(self.scan, "tramba_order", None)
        if (not self.fused or order is None or getattr(self.merge, "tramba_order", None) != order
                or self.selective_scan is not ops.SelectiveScanOflex or not x.is_cuda):
            return None
        B, D, H, W = x.shape
        K, N = self.k_group, self.A_logs.shape[1]
        if not fused.supported(order, K, D, H, W, N, x.dtype, need_bwd=torch.is_grad_enabled()):
            return None
        return order, self.scan.tramba_param(H)

    @property
    def _act_in_norm(self):
        """channel-first: out_norm, the cast and out_act (vmamba.py:271,287) are one LayerNorm2d(+GELU) kernel."""
        return self.channel_first and isinstance(self.out_act, nn.GELU) and self.out_act.approximate == "none"

    def _out_norm_act(self, y, dtype):
        if self._act_in_norm:
            return self.out_norm(y, act="gelu", to_dtype=dtype)
        return self.out_norm(y).to(dtype)

    def forward_core_fused(self, x, order, param):
        """Same function as forward_core (Models/vmamba.py:231-273).  x_proj is a per-token linear map,
        so it commutes with the cross-scan gather: it is evaluated ONCE on the image-ordered map (one
        dense 1x1 conv instead of K grouped ones on the K-times larger xs) and its R+2N channels -- 1/25
        of the feature map -- are gathered into scan order; dt_proj then yields delta in scan order
        exactly as in the reference, and the fused kernel applies the scan order to x and to the merged
        output through its shared-memory addressing (xs / ys never exist)."""
        B, D, H, W = x.shape
        K, D, R = self.dt_projs_weight.shape
        N = self.A_logs.shape[1]
        L = H * W
        xf = x.contiguous().view(B, D, L)
        if _PROJ_GEMM:
            x_dbl = fused.ProjX.apply(self.x_proj_weight.view(-1, D), xf).view(B, K, R + 2 * N, L)
        else:
            x_dbl = F.conv1d(xf, self.x_proj_weight.view(-1, D, 1), bias=None).view(B, K, R + 2 * N, L)
        if _PROJ_KERNEL:    # one kernel: image -> scan order + split (contiguous dts_r, Bs, Cs)
            dts, Bs, Cs = fused.proj_gather(x_dbl, order, param, H, R, N)
        else:
            tok = fused.token_index(order, H, param, x.device)                       # (K, L)
            x_dbl = x_dbl.gather(3, tok.view(1, K, 1, L).expand(B, K, R + 2 * N, L))  # image -> scan order
            dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
        if _PROJ_GEMM:
            dts = fused.ProjDt.apply(self.dt_projs_weight, dts)                  # (K, D, R) @ (B, K, R, L)
        else:
            dts = F.conv1d(dts.contiguous().view(B, -1, L), self.dt_projs_weight.view(K * D, -1, 1), groups=K)
        As = -torch.exp(self.A_logs.to(torch.float)).view(-1)
        y = fused.ss2d_core_fused(x, dts.view(B, K, D, L), Bs.reshape(B, K, L), Cs.reshape(B, K, L), As,
                                  self.Ds.to(torch.float), self.dt_projs_bias.view(-1).to(torch.float), order, param)
        if self.channel_first:
            return self._out_norm_act(y.view(B, -1, H, W), x.dtype)
        return self.out_norm(y.transpose(1, 2).contiguous().view(B, H, W, -1)).to(x.dtype)

    def forward_core(self, x):
        """Models/vmamba.py:231-273 (force_fp32=False, no_einsum=True, ssoflex=True)."""
        route = self._fused_order(x)
        if route is not None:
            return self.forward_core_fused(x, *route)
        B, D, H, W = x.shape
        D, N = self.A_logs.shape
        K, D, R = self.dt_projs_weight.shape
        L = H * W
        xs = self.scan.apply(x)                                                        # (B, K, D, L)
        x_dbl = F.conv1d(xs.view(B, -1, L), self.x_proj_weight.view(-1, D, 1), bias=None, groups=K)
        dts, Bs, Cs = torch.split(x_dbl.view(B, K, -1, L), [R, N, N], dim=2)
        dts = F.conv1d(dts.contiguous().view(B, -1, L), self.dt_projs_weight.view(K * D, -1, 1), groups=K)
        xs = xs.view(B, -1, L)
        dts = dts.contiguous().view(B, -1, L)
        As = -torch.exp(self.A_logs.to(torch.float))
        Bs = Bs.contiguous().view(B, K, N, L)
        Cs = Cs.contiguous().view(B, K, N, L)
        Ds = self.Ds.to(torch.float)
        delta_bias = self.dt_projs_bias.view(-1).to(torch.float)
        ys = self.selective_scan.apply(xs, dts, As, Bs, Cs, Ds, delta_bias, True, -1, -1, True)
        y = self.merge.apply(ys.view(B, K, -1, H, W))                                   # (B, D, L)
        y = y.view(B, -1, H, W)
        if self.channel_first:
            return self._out_norm_act(y, x.dtype)
        y = y.view(B, -1, H * W).transpose(1, 2).contiguous().view(B, H, W, -1)
        y = self.out_norm(y)
        return y.to(x.dtype)

    def forward(self, x, **kwargs):
        """Models/vmamba.py:275-291."""
        x = self.in_proj(x)
        z = None
        if not self.disable_z:
            x, z = x.chunk(2, dim=(1 if self.channel_first else -1))
            z = self.act(z)
        if not self.channel_first:
            x = x.permute(0, 3, 1, 2).contiguous()
        if (self.with_dconv and self.fused_dwconv and isinstance(self.act, nn.SiLU)
                and dwconv.supported(x, self.conv2d.weight)):
            x = dwconv.dwconv2d(x, self.conv2d.weight, self.conv2d.bias, "silu")   # conv2d + act, one kernel
        else:
            if self.with_dconv:
                x = self.conv2d(x)
            x = self.act(x)
        y = self.forward_core(x)
        if not self._act_in_norm:
            y = self.out_act(y)
        if not self.disable_z:
            y = y * z
        return self.dropout(self.out_proj(y))
