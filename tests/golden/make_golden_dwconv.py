"""Golden fixture for the depthwise conv + SiLU of SS2D.forward: runs the reference's own ops --
nn.Conv2d(C, C, k, padding=(k-1)//2, groups=C) and nn.SiLU, exactly as constructed at
Models/vmamba.py:72-79 and applied at :281-285 -- on CPU (torch), forward and backward.

    python tests/golden/make_golden_dwconv.py    ->  tests/golden/dwconv.npz
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ((2, 5, 12, 12, 3, True), (1, 4, 24, 24, 3, False), (2, 3, 10, 14, 5, True), (1, 2, 17, 9, 7, True))


def main():
    out = {}
    for i, (B, C, H, W, k, has_bias) in enumerate(CASES):
        rs = np.random.RandomState(4000 + i)
        x = rs.standard_normal((B, C, H, W)).astype(np.float32)
        w = (0.5 * rs.standard_normal((C, 1, k, k))).astype(np.float32)
        b = (0.3 * rs.standard_normal(C)).astype(np.float32)
        gy = rs.standard_normal((B, C, H, W)).astype(np.float32)
        for act in ("none", "silu"):
            conv = torch.nn.Conv2d(C, C, k, padding=(k - 1) // 2, groups=C, bias=has_bias)
            with torch.no_grad():
                conv.weight.copy_(torch.from_numpy(w))
                if has_bias:
                    conv.bias.copy_(torch.from_numpy(b))
            xt = torch.from_numpy(x).requires_grad_(True)
            y = conv(xt)
            if act == "silu":
                y = torch.nn.SiLU()(y)
            params = [xt, conv.weight] + ([conv.bias] if has_bias else [])
            grads = torch.autograd.grad(y, params, torch.from_numpy(gy))
            out[f"dw_{i}_{act}_y"] = y.detach().numpy()
            out[f"dw_{i}_{act}_gx"] = grads[0].numpy()
            out[f"dw_{i}_{act}_gw"] = grads[1].numpy()
            if has_bias:
                out[f"dw_{i}_{act}_gb"] = grads[2].numpy()
    np.savez_compressed(os.path.join(HERE, "dwconv.npz"), **out)
    print("dwconv.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
