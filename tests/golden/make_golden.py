"""Generate the committed golden fixtures by running the REAL reference (from
/root/reference, CPU, under tests/golden/_ref_shim.py) in the build container.

    python tests/golden/make_golden.py

Writes (all small, compressed):
  orders.npz     flat scan-order tables from the reference generators
                 (generate_dilation_indices / generate_window_indices / generate_indices)
  scanmerge.npz  CrossScan*/CrossMerge* forward + backward outputs on seeded inputs
  dct.npz        DCT2D (high, low) on seeded inputs
Inputs are produced with numpy RandomState so the tests can rebuild them without torch RNG.
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
    from Models.SS2D.Dilation import generate_dilation_indices  # noqa: E402
    from Models.SS2D.Window import generate_window_indices  # noqa: E402
    from Models.SS2D.SpiralLine import generate_indices  # noqa: E402
    from Models.DCT_2D import DCT2D  # noqa: E402


def flat_rowmajor(tabs, H):  # Dilation.py:66 / Window.py:56
    return np.stack([(t[:, 0] * H + t[:, 1]).numpy() for t in tabs]).astype(np.int32)


def flat_line(tabs, H):  # SpiralLine.py:103
    return np.stack([(t[:, 0] + t[:, 1] * H).numpy() for t in tabs]).astype(np.int32)


def orders():
    out = {}
    ws = {12: 4, 24: 8, 48: 12, 96: 16, 192: 32}      # csms6s.py:107-108 (+ build-side 192)
    for H in (12, 24, 48, 96, 192):
        out[f"dilation_{H}_4"] = flat_rowmajor(generate_dilation_indices(H, H, 4), H)
        out[f"window_{H}_{ws[H]}"] = flat_rowmajor(generate_window_indices(H, H, ws[H]), H)
    out["dilation_4_3"] = flat_rowmajor(generate_dilation_indices(4, 4, 3), 4)  # Dilation.py:103
    out["window_8_4"] = flat_rowmajor(generate_window_indices(8, 8, 4), 8)      # Window.py:93
    for H in (6, 7, 12, 14, 24, 28, 48, 56, 96, 192):
        out[f"line_{H}"] = flat_line(generate_indices(H, H), H)
    # the tables the reference actually holds at import (csms6s.py:60,109,158)
    for H in (12, 24, 48, 96):
        assert np.array_equal(flat_rowmajor(csms6s.dilation_index[str(H)], H), out[f"dilation_{H}_4"])
        assert np.array_equal(flat_rowmajor(csms6s.window_index[str(H)], H), out[f"window_{H}_{ws[H]}"])
        assert np.array_equal(flat_line(csms6s.spiral_line_index[str(H)], H), out[f"line_{H}"])
    np.savez_compressed(os.path.join(HERE, "orders.npz"), **out)
    print("orders.npz", {k: v.shape for k, v in out.items()})


PAIRS = {
    "std": (csms6s.CrossScan, csms6s.CrossMerge, 4),
    "dilation": (csms6s.CrossScan_Dilation, csms6s.CrossMerge_Dilation, 4),
    "window": (csms6s.CrossScan_Window, csms6s.CrossMerge_Window, 4),
    "line": (csms6s.CrossScan_Line, csms6s.CrossMerge_Line, 8),
}


def scanmerge():
    out = {}
    for H in (12, 24):
        B, C = 2, 3
        for fam, (S, M, K) in PAIRS.items():
            rs = np.random.RandomState(1000 + H)
            x = rs.standard_normal((B, C, H, H)).astype(np.float32)
            ys = rs.standard_normal((B, K, C, H, H)).astype(np.float32)
            xt = torch.from_numpy(x).requires_grad_(True)
            xs = S.apply(xt)                                   # (B,K,C,L)
            g = torch.from_numpy(ys).view(B, K, C, H * H)
            (gx,) = torch.autograd.grad(xs, xt, g)             # CrossScan.backward == merge
            yt = torch.from_numpy(ys).requires_grad_(True)
            y = M.apply(yt)                                    # (B,C,L)
            gy = torch.from_numpy(x).view(B, C, H * H)
            (gys,) = torch.autograd.grad(y, yt, gy)            # CrossMerge.backward == scan
            out[f"{fam}_{H}_scan"] = xs.detach().numpy()
            out[f"{fam}_{H}_scan_bwd"] = gx.numpy()
            out[f"{fam}_{H}_merge"] = y.detach().numpy()
            out[f"{fam}_{H}_merge_bwd"] = gys.numpy()
    np.savez_compressed(os.path.join(HERE, "scanmerge.npz"), **out)
    print("scanmerge.npz", len(out), "arrays")


def dct():
    out = {}
    for (B, C, H) in ((2, 3, 24), (1, 2, 48), (1, 1, 96), (2, 2, 12)):
        rs = np.random.RandomState(2000 + H)
        x = rs.standard_normal((B, C, H, H)).astype(np.float32)
        m = DCT2D(H, H)
        with torch.no_grad():
            high, low = m(torch.from_numpy(x))
        out[f"dct_{B}_{C}_{H}_high"] = high.numpy()
        out[f"dct_{B}_{C}_{H}_low"] = low.numpy()
        out[f"dct_{H}_wx"] = m.dct_x.weight.numpy()
        out[f"dct_{H}_wy"] = m.dct_y.weight.numpy()
    np.savez_compressed(os.path.join(HERE, "dct.npz"), **out)
    print("dct.npz", len(out), "arrays")


if __name__ == "__main__":
    orders()
    scanmerge()
    dct()
