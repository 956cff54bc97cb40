"""Golden fixtures for the reference's ABLATION scan orders, made by running the
REAL reference classes (Models/SS2D/csms6s.py:219-737, CPU, under _ref_shim) in the build container.

    python tests/golden/make_golden_ablation.py        -> tests/golden/ablation_orders.npz

  table_<fam>_<H>        (K, L) uint16: CrossScan_<fam>.apply(arange image) -- the gather table itself
  <fam>_<H>_scan / _scan_bwd / _merge / _merge_bwd      forward + backward outputs on seeded inputs
Inputs come from numpy RandomState so the tests rebuild them without torch RNG.
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
    from Models.SS2D import csms6s  # noqa: E402

PAIRS = {
    "line4": (csms6s.CrossScan_Line_4direction, csms6s.CrossMerge_Line_4direction, 4),
    "spiral": (csms6s.CrossScan_Spiral, csms6s.CrossMerge_Spiral, 8),
    "hilbert": (csms6s.CrossScan_Hilbert, csms6s.CrossMerge_Hilbert, 4),
    "diagonal": (csms6s.CrossScan_Diagonal, csms6s.CrossMerge_Diagonal, 8),
    "ds": (csms6s.CrossScan_DS, csms6s.CrossMerge_DS, 8),
    "ab2": (csms6s.CrossScan_Ab_2direction, csms6s.CrossMerge_Ab_2direction, 4),
    "ab1": (csms6s.CrossScan_Ab_1direction, csms6s.CrossMerge_Ab_1direction, 4),
}


def main():
    out = {}
    for H in (12, 24, 48, 96):
        L = H * H
        x = torch.arange(L, dtype=torch.float32).view(1, 1, H, H)      # exact in fp32 (L <= 9216)
        for fam, (S, M, K) in PAIRS.items():
            tab = S.apply(x)[0, :, 0].numpy()
            assert tab.shape == (K, L) and np.array_equal(tab, np.round(tab))
            out[f"table_{fam}_{H}"] = tab.astype(np.uint16)
    H, B, C = 12, 2, 3
    for fam, (S, M, K) in PAIRS.items():
        rs = np.random.RandomState(3000 + K)
        x = rs.standard_normal((B, C, H, H)).astype(np.float32)
        ys = rs.standard_normal((B, K, C, H, H)).astype(np.float32)
        xt = torch.from_numpy(x).requires_grad_(True)
        xs = S.apply(xt)
        (gx,) = torch.autograd.grad(xs, xt, torch.from_numpy(ys).view(B, K, C, H * H))
        yt = torch.from_numpy(ys).requires_grad_(True)
        y = M.apply(yt)
        (gys,) = torch.autograd.grad(y, yt, torch.from_numpy(x).view(B, C, H * H))
        out[f"{fam}_{H}_scan"] = xs.detach().numpy()
        out[f"{fam}_{H}_scan_bwd"] = gx.numpy().reshape(B, C, H * H)
        out[f"{fam}_{H}_merge"] = y.detach().numpy().reshape(B, C, H * H)
        out[f"{fam}_{H}_merge_bwd"] = gys.numpy().reshape(B, K, C, H * H)
    path = os.path.join(HERE, "ablation_orders.npz")
    np.savez_compressed(path, **out)
    print(path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
