"""Generate tests/golden/model.npz by running the REAL reference model (Trambav6.bulid_model,
/root/reference, CPU, under _ref_shim) with the oracle selective scan injected at the
extension boundary (Models.vmamba.SelectiveScanOflex).

    python tests/golden/make_golden_model.py

Config: Tramba-V-TSOD, img 384, dims=16 (channel width /8; depths, scan orders, map sizes and
sequence lengths are the full model's), eval mode, deterministic weights (tests/_det_weights.py).
Stores the 4 deep-supervision outputs (the 384^2 head strided by 4), the train.py loss and a few
parameter gradients.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref_shim  # noqa: E402

_ref_shim.install()
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from _det_weights import det_fill, det_input  # noqa: E402
from _oracle_ops import OracleSelectiveScan  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    import Models.vmamba as vm  # noqa: E402
    vm.SelectiveScanOflex = OracleSelectiveScan
    import Trambav6  # noqa: E402
    from utils.loss import iou_loss  # noqa: E402

GRAD_KEYS = ["vssm_encoder.patch_embed.0.weight", "vssm_encoder.layers.0.blocks.0.op.A_logs",
             "vssm_encoder.layers.2.blocks.7.op.dt_projs_bias", "decoder.stage_layers.2.blocks.1.op.Ds",
             "decoder.guide_layers.1.attn.h_ssm.x_proj_weight", "decoder.guide_layers.2.attn.l_ssm.dt_projs_weight",
             "decoder.seg_layers.3.weight"]


def main():
    dims = 16
    model = Trambav6.bulid_model(deep_supervision=True, use_pretrain=False, img_size=384, dims=dims).eval()
    det_fill(model)
    x = det_input((1, 3, 384, 384), 7)
    label = (det_input((1, 1, 384, 384), 8) > 0).float()
    outs = model(x)
    ups = [o if o.shape[-1] == 384 else F.interpolate(o, (384, 384), mode="bilinear") for o in outs]
    loss = sum(F.binary_cross_entropy_with_logits(o, label) + iou_loss(o, label) for o in ups)  # train.py:74-85
    loss.backward()
    params = dict(model.named_parameters())
    out = {"dims": np.int32(dims), "loss": np.float64(loss.item())}
    for i, o in enumerate(outs):
        a = o.detach().numpy()
        out[f"out{i}"] = a if a.shape[-1] < 384 else a[..., ::4, ::4]
    for k in GRAD_KEYS:
        out["grad:" + k] = params[k].grad.numpy()
    np.savez_compressed(os.path.join(HERE, "model.npz"), **out)
    print("model.npz", {k: getattr(v, "shape", v) for k, v in out.items()}, "loss", loss.item())


if __name__ == "__main__":
    main()
