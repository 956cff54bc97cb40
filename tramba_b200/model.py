"""Tramba-V (VMamba-B encoder + Dual-Frequency / Helix decoder) -- the CALLER of the hot path.

This file is the benchmark / parity harness for BASELINE.json configs 1, 3, 4 and 5: everything
in it is plain PyTorch (convolutions, LayerNorm, MLPs -- the ops the north star leaves alone);
the only product code it reaches is tramba_b200.SS2D (33 calls per image) and
tramba_b200.DCT2D (3 calls per image).  It restates the reference's module tree so that
  * `state_dict()` keys and shapes equal those of `Trambav6.bulid_model(...)`
    (Trambav6.py:168-200) -- a reference checkpoint loads with strict=True, and
  * the forward pass is the same function (tests/test_model_parity.py checks both against the
    real reference in the build container, and against a committed golden fixture elsewhere).

Reference map:
    VSSBlock                 Models/vmamba.py:327-396      Mlp                Models/modules.py:134-153
    VSSMEncoder              Models/vmamba.py:399-518      PatchExpand        Models/modules.py:183-221
    DWMSMlp / MultiScale..   Models/vmamba.py:603-704      FinalPatchExpand   Models/modules.py:224-250
    FreqSS2Dv6 / FreqBlockv6 Models/freq_mamba.py:11-82    FreqExpand2D       Models/modules.py:687-696
    VSSMDecoder / BaseUMamba Trambav6.py:13-165            loss               train.py:74-85, utils/loss.py:6-11
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dwconv, ops
from .dct import DCT2D
from .ss2d import SS2D, LayerNorm2d, Linear2d


class DropPath(nn.Module):
    """Per-sample stochastic depth, scaled by the keep probability (timm.models.layers.DropPath)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return x * mask


def _gemm_dtype():
    """dtype a LayerNorm2d may emit directly when its only consumer is an autocast GEMM (the reference's
    F.layer_norm returns fp32 under autocast and the GEMM casts it; same values, one pass less)."""
    return torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else None


def _pixel_unshuffle_cf(x, p):
    """'b (p1 p2 c) h w -> b c (h p1) (w p2)' (modules.py:213,247,692)."""
    B, C, H, W = x.shape
    c = C // (p * p)
    return x.view(B, p, p, c, H, W).permute(0, 3, 4, 1, 5, 2).reshape(B, c, H * p, W * p)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = Linear2d(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = Linear2d(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _DWConv(nn.Module):
    def __init__(self, ch, k):
        super().__init__()
        self.dw_conv = nn.Conv2d(ch, ch, k, stride=1, padding=(k - 1) // 2, bias=True, groups=ch)

    def forward(self, x):
        return self.dw_conv(x)


class DWMSMlp(nn.Module):
    """Multi-scale depthwise MLP of the Helix decoder block (vmamba.py:603-629)."""

    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = Linear2d(dim, hidden)
        self.act = nn.GELU()
        self.dwc3, self.dwc5, self.dwc7 = _DWConv(hidden, 3), _DWConv(hidden, 5), _DWConv(hidden, 7)
        self.fc2 = Linear2d(hidden, dim)
        ident = torch.zeros(1, 1, 7, 7)
        ident[0, 0, 3, 3] = 1.0
        self.register_buffer("_ident", ident, persistent=False)     # not in the state_dict (reference keys only)

    fused = os.environ.get("TRAMBA_DWCONV", "fused") != "torch"

    def _taps(self):
        """x + dwc3(x) + dwc5(x) + dwc7(x) is ONE 7x7 depthwise conv: the zero-padded 3x3 and 5x5 taps and
        the identity add into the 7x7 taps (same 'same' padding), the biases add."""
        w7 = self.dwc7.dw_conv.weight
        w = w7 + F.pad(self.dwc5.dw_conv.weight, (1, 1, 1, 1)) + F.pad(self.dwc3.dw_conv.weight, (2, 2, 2, 2))
        return w + self._ident, self.dwc3.dw_conv.bias + self.dwc5.dw_conv.bias + self.dwc7.dw_conv.bias

    def forward(self, x):
        x = self.fc1(x)
        if self.fused and dwconv.supported(x, self.dwc7.dw_conv.weight):
            w, b = self._taps()
            x = dwconv.dwconv2d(x, w, b, "none")
        else:
            x = x + self.dwc3(x) + self.dwc5(x) + self.dwc7(x)
        return self.fc2(self.act(x))


def _ss2d(dim, scan, merge, k_group):
    return SS2D(d_model=dim, d_state=1, ssm_ratio=2.0, dt_rank="auto", act_layer=nn.SiLU, d_conv=3,
                conv_bias=False, dropout=0.0, initialize="v0", channel_first=True, bias=False, disable_z=True,
                scan=scan, merge=merge, k_group=k_group)


class VSSBlock(nn.Module):
    """x + DropPath(SS2D(LN(x))); x + DropPath(Mlp(LN(x)))  (vmamba.py:384-396)."""

    def __init__(self, dim, drop_path):
        super().__init__()
        self.norm = LayerNorm2d(dim)
        self.op = _ss2d(dim, ops.CrossScan, ops.CrossMerge, 4)
        self.drop_path = DropPath(drop_path)
        self.norm2 = LayerNorm2d(dim)
        self.mlp = Mlp(dim, int(dim * 4.0))

    def forward(self, x):
        x = x + self.drop_path(self.op(self.norm(x, to_dtype=_gemm_dtype())))
        return x + self.drop_path(self.mlp(self.norm2(x, to_dtype=_gemm_dtype())))


class MultiScaleDecoderBlock(nn.Module):
    """Helix-SS2D block: K = 8 line scan + DWMSMlp (vmamba.py:632-704)."""

    def __init__(self, dim, drop_path):
        super().__init__()
        self.norm1 = LayerNorm2d(dim)
        self.op = _ss2d(dim, ops.CrossScan_Line, ops.CrossMerge_Line, 8)
        self.drop_path = DropPath(drop_path)
        self.norm2 = LayerNorm2d(dim)
        self.mlp = DWMSMlp(dim, int(dim * 4.0))

    def forward(self, x):
        x = x + self.drop_path(self.op(self.norm1(x, to_dtype=_gemm_dtype())))
        return x + self.drop_path(self.mlp(self.norm2(x, to_dtype=_gemm_dtype())))


def _stage(block, dim, drop_paths):
    return nn.Sequential(OrderedDict(blocks=nn.Sequential(*[block(dim, dp) for dp in drop_paths])))


def _init_weights(m):
    """VSSMEncoder._init_weights / VSSMDecoder._init_weights (vmamba.py:459-471, Trambav6.py:86-98)."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=0.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight, a=1e-2)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)


class VSSMEncoder(nn.Module):
    def __init__(self, in_chans=3, depths=(2, 2, 15, 2), dims=128, drop_path_rate=0.6):
        super().__init__()
        dims = [int(dims * 2 ** i) for i in range(len(depths))]
        self.dims = dims
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        e = dims[0]
        # patch 4 = two stride-2 3x3 convs (vmamba.py:473-489); Identity slots keep the key indices
        self.patch_embed = nn.Sequential(
            nn.Conv2d(in_chans, e // 2, kernel_size=3, stride=2, padding=1), nn.Identity(), LayerNorm2d(e // 2),
            nn.Identity(), nn.GELU(), nn.Conv2d(e // 2, e, kernel_size=3, stride=2, padding=1), nn.Identity(),
            LayerNorm2d(e))
        self.layers = nn.ModuleList()
        self.downsample = nn.ModuleList()
        for i, depth in enumerate(depths):
            self.downsample.append(nn.Sequential(
                nn.Identity(), nn.Conv2d(dims[i], dims[i] * 2, kernel_size=3, stride=2, padding=1), nn.Identity(),
                LayerNorm2d(dims[i] * 2)) if i < len(depths) - 1 else nn.Identity())
            self.layers.append(_stage(VSSBlock, dims[i], dpr[sum(depths[:i]):sum(depths[:i + 1])]))
        self.apply(_init_weights)

    def forward(self, x):
        out = [x]
        x = self.patch_embed(x)
        for s, layer in enumerate(self.layers):
            x = layer(x)
            out.append(x)
            x = self.downsample[s](x)
        return out


class _Expand(nn.Module):
    """PatchExpand (x2), FinalPatchExpand_X4 (x4) and FreqExpand2D (x2, channel-preserving)."""

    def __init__(self, dim, out_mult, scale, norm_dim):
        super().__init__()
        self.scale = scale
        self.expand = Linear2d(dim, out_mult * dim, bias=False)
        self.norm = LayerNorm2d(norm_dim)

    def forward(self, x):
        return self.norm(_pixel_unshuffle_cf(self.expand(x), self.scale))


class FreqSS2Dv6(nn.Module):
    """Dual-Frequency VSS: DCT split -> expand -> window / dilated SS2D -> sigmoid gate
    (freq_mamba.py:11-57).  Note DCT2D returns (high, low)."""

    def __init__(self, dim, input_resolution):
        super().__init__()
        self.DCT2D = DCT2D(input_resolution[0], input_resolution[1])
        self.l_expand = _Expand(dim, 4, 2, dim)
        self.l_ssm = _ss2d(dim, ops.CrossScan_Dilation, ops.CrossMerge_Dilation, 4)
        self.h_expand = _Expand(dim, 4, 2, dim)
        self.h_ssm = _ss2d(dim, ops.CrossScan_Window, ops.CrossMerge_Window, 4)
        self.concat_back_dim = Linear2d(dim * 2, dim, bias=False)

    def forward(self, x):
        high, low = self.DCT2D(x)
        hifi = self.h_ssm(self.h_expand(high))
        lofi = self.l_ssm(self.l_expand(low))
        return torch.sigmoid(self.concat_back_dim(torch.cat((hifi, lofi), 1))) * x


class FreqBlockv6(nn.Module):
    def __init__(self, dim, input_resolution):
        super().__init__()
        self.norm1 = LayerNorm2d(dim)
        self.attn = FreqSS2Dv6(dim, input_resolution)
        self.norm2 = LayerNorm2d(dim)
        self.mlp = Mlp(dim, int(dim * 4.0))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x, to_dtype=_gemm_dtype()))


class VSSMDecoder(nn.Module):
    def __init__(self, deep_supervision, features_per_stage, drop_path_rate=0.0, depths=(2, 2, 2, 2), img_size=384):
        super().__init__()
        ch = list(features_per_stage)
        n = len(ch)
        self.deep_supervision = deep_supervision
        dpr = [x.item() for x in torch.linspace(drop_path_rate, 0, (n - 1) * 2)]
        res = img_size // 2 ** len(depths)  # Trambav6.py:31: deepest skip is img/32 ... as written: img/16
        self.stage_layers = nn.ModuleList()
        self.expand_layers = nn.ModuleList()
        self.guide_layers = nn.ModuleList()
        self.seg_layers = nn.ModuleList()
        self.concat_back_dim = nn.ModuleList()
        for stage in range(1, n):
            below, skip = ch[-stage], ch[-(stage + 1)]
            self.expand_layers.append(_Expand(below, 2, 2, below // 2))
            r = res * 2 ** (stage - 1)
            self.guide_layers.append(FreqBlockv6(skip, (r, r)))
            self.stage_layers.append(_stage(MultiScaleDecoderBlock, skip,
                                            dpr[sum(depths[:stage - 1]):sum(depths[:stage])]))
            self.seg_layers.append(nn.Conv2d(skip, 1, 1, 1, 0, bias=True))
            self.concat_back_dim.append(Linear2d(2 * skip, skip))
        self.expand_layers.append(_Expand(ch[0], 16, 4, ch[0]))
        self.stage_layers.append(nn.Identity())
        self.seg_layers.append(nn.Conv2d(ch[0], 1, 1, 1, 0, bias=True))
        self.apply(_init_weights)

    def forward(self, skips):
        x_low = skips[-1]
        outs = []
        last = len(self.stage_layers) - 1
        for s in range(len(self.stage_layers)):
            x = self.expand_layers[s](x_low)
            if s < last:
                mid = self.guide_layers[s](skips[-(s + 2)])
                x = self.concat_back_dim[s](torch.cat((x, mid), 1))
            x = self.stage_layers[s](x)
            if self.deep_supervision or s == last:
                outs.append(self.seg_layers[s if self.deep_supervision else -1](x))
            x_low = x
        return outs if self.deep_supervision else outs[0]


class TrambaV(nn.Module):
    """BaseUMamba (Trambav6.py:142-154): submodule names vssm_encoder / decoder as in the reference."""

    def __init__(self, deep_supervision=True, img_size=384, dims=128, depths=(2, 2, 2, 2), enc_depths=(2, 2, 15, 2),
                 in_chans=3):
        super().__init__()
        self.vssm_encoder = VSSMEncoder(in_chans=in_chans, depths=enc_depths, dims=dims, drop_path_rate=0.6)
        self.decoder = VSSMDecoder(deep_supervision, [dims, dims * 2, dims * 4, dims * 8], drop_path_rate=0.2,
                                   depths=depths, img_size=img_size)

    def forward(self, x):
        return self.decoder(self.vssm_encoder(x))


def build_model(deep_supervision=True, img_size=384, dims=128, depths=(2, 2, 2, 2), **kw):
    """Signature of Trambav6.bulid_model (Trambav6.py:168-200) minus the checkpoint arguments."""
    return TrambaV(deep_supervision=deep_supervision, img_size=img_size, dims=dims, depths=depths, **kw)


def iou_loss(pred, mask):
    """utils/loss.py:6-11."""
    pred = torch.sigmoid(pred)
    inter = (pred * mask).sum(dim=(2, 3))
    union = (pred + mask).sum(dim=(2, 3))
    return (1 - (inter + 1) / (union - inter + 1)).mean()


def tsod_loss(outputs, label):
    """Deep-supervision loss of train.py:74-85: sum over the 4 heads (bilinearly upsampled to the
    label size) of BCE-with-logits + IoU loss, unit weights."""
    size = label.shape[-2:]
    total = 0.0
    for o in outputs:
        o = o.float()
        if o.shape[-2:] != size:
            o = F.interpolate(o, size, mode="bilinear")
        total = total + F.binary_cross_entropy_with_logits(o, label) + iou_loss(o, label)
    return total
