"""Golden fixture for LayerNorm2d (+ GELU): runs the REAL reference class
(Models/modules.py:22-27 LayerNorm2d, torch.nn.GELU as used at Models/vmamba.py:102) on CPU under
tests/golden/_ref_shim.py in the build container.

    python tests/golden/make_golden_layernorm.py    ->  tests/golden/layernorm2d.npz
Inputs come from numpy RandomState so the tests can rebuild them.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_shim  # noqa: E402

_ref_shim.install()
import torch  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    from Models.modules import LayerNorm2d  # noqa: E402

CASES = ((2, 8, 6, 6), (1, 64, 12, 12), (2, 256, 5, 7), (1, 130, 3, 11))


def main():
    out = {}
    for i, (B, C, H, W) in enumerate(CASES):
        rs = np.random.RandomState(3000 + i)
        x = (rs.standard_normal((B, C, H, W)) * 2.0 + 0.7).astype(np.float32)
        w = (1.0 + 0.3 * rs.standard_normal(C)).astype(np.float32)
        b = (0.2 * rs.standard_normal(C)).astype(np.float32)
        gy = rs.standard_normal((B, C, H, W)).astype(np.float32)
        for act in ("none", "gelu"):
            m = LayerNorm2d(C)
            with torch.no_grad():
                m.weight.copy_(torch.from_numpy(w)), m.bias.copy_(torch.from_numpy(b))
            xt = torch.from_numpy(x).requires_grad_(True)
            y = m(xt)
            if act == "gelu":
                y = torch.nn.GELU()(y)
            gx, gw, gb = torch.autograd.grad(y, [xt, m.weight, m.bias], torch.from_numpy(gy))
            out[f"ln_{i}_{act}_y"] = y.detach().numpy()
            out[f"ln_{i}_{act}_gx"] = gx.numpy()
            out[f"ln_{i}_{act}_gw"] = gw.numpy()
            out[f"ln_{i}_{act}_gb"] = gb.numpy()
    np.savez_compressed(os.path.join(HERE, "layernorm2d.npz"), **out)
    print("layernorm2d.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
