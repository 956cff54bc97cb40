"""Generates tests/golden/dct192.npz from the REAL reference module (Models/DCT_2D.py:6-77, imported from
/root/reference under the CPU shim) at H = W = 192 -- the feature size of the 768^2 stress configuration (BASELINE
config 5), which the library serves with its streamed large-plane DCT kernel (csrc/dct2d_tc.cu).  The bases are pinned through a checksum
and a strided sample (the full 192 x 192 matrices would dominate the fixture); the outputs are stored in full.  Also a forward + autograd-backward pair for the adjoint kernel.
    python tests/golden/make_golden_dct192.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_shim  # noqa: E402

_ref_shim.install()
sys.path.insert(0, _ref_shim.REF)
from Models.DCT_2D import DCT2D  # noqa: E402


def main():
    H = 192
    rs = np.random.RandomState(2000 + H)
    x = rs.standard_normal((1, 2, H, H)).astype(np.float32)
    m = DCT2D(H, H)
    xt = torch.from_numpy(x).requires_grad_(True)
    high, low = m(xt)
    gh = torch.from_numpy(rs.standard_normal(tuple(high.shape)).astype(np.float32))
    gl = torch.from_numpy(rs.standard_normal(tuple(low.shape)).astype(np.float32))
    (gx,) = torch.autograd.grad([high, low], xt, [gh, gl])
    wx, wy = m.dct_x.weight.numpy(), m.dct_y.weight.numpy()
    out = {"high": high.detach().numpy(), "low": low.detach().numpy(), "gh": gh.numpy(), "gl": gl.numpy(),
           "gx": gx.numpy(), "wx_sample": wx[::7, ::5].copy(), "wy_sample": wy[::7, ::5].copy(),
           "wx_sum": np.float64(wx.astype(np.float64).sum()), "wy_abs_sum": np.float64(np.abs(wy.astype(np.float64)).sum())}
    np.savez_compressed(os.path.join(HERE, "dct192.npz"), **out)
    print("dct192.npz", {k: getattr(v, "shape", ()) for k, v in out.items()})


if __name__ == "__main__":
    main()
