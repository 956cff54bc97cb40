"""CPU ORACLE -- test infrastructure only, NOT product code.

A CPU restatement of the reference's 2D selective-scan hot path (scan orders, cross-scan /
cross-merge, selective-scan recurrence fwd/bwd, DCT2D split).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it; `tramba_b200/` never does (tests/test_boundary.py greps for that).

Pinning status:
  scan orders / cross-scan / cross-merge : PINNED bit-exact to the reference generators
                                           (tests/golden/orders.npz, scanmerge.npz)
  DCT2D                                  : PINNED to the reference module run on CPU
                                           (tests/golden/dct.npz) and scipy.fft.dctn
  selective-scan recurrence              : PARITY UNPINNED against the reference's own code -- the
                                           recurrence is an external CUDA extension absent from
                                           /root/reference; restated from the published
                                           selective_scan_ref algorithm and PINNED to an independent
                                           executed implementation of it (HF transformers
                                           MambaMixer.slow_forward: tests/golden/sscan_hf.npz,
                                           forward + all gradients, fp64)
  LayerNorm2d (+GELU)                    : PINNED to the reference class run on CPU
                                           (tests/golden/layernorm2d.npz)
  depthwise conv (+SiLU)                 : PINNED to the reference's own torch ops (nn.Conv2d
                                           groups=C + nn.SiLU) run on CPU (tests/golden/dwconv.npz)
"""
