"""INTEGRATION.md, executed: the REAL reference model (Trambav6.bulid_model from /root/reference, unmodified, under
tests/golden/_ref_shim.py) built on top of the product's Python host at the three integration depths, and compared
with tests/golden/model.npz (the same reference model with the oracle scan, tests/golden/make_golden_model.py).

No GPU here, so the C-ABI compute entry points are emulated on host memory (tests/_abi_emulator.py); everything
above the C-ABI is the shipped code: the `selective_scan_cuda_oflex` module, the CrossScan* / CrossMerge* /
SelectiveScanOflex classes, tramba_b200.SS2D and tramba_b200.DCT2D inside the reference's own module tree.
Skipped where /root/reference does not exist (the GPU box)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import _ref_shim  # noqa: E402

pytestmark = pytest.mark.skipif(not _ref_shim.available(), reason="/root/reference not present")

import tramba_b200 as tb  # noqa: E402
from _abi_emulator import installed  # noqa: E402
from _det_weights import det_fill, det_input  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "model.npz"))
OPS = ("CrossScan", "CrossMerge", "CrossScan_Dilation", "CrossMerge_Dilation", "CrossScan_Window",
       "CrossMerge_Window", "CrossScan_Line", "CrossMerge_Line", "SelectiveScanOflex")


@pytest.fixture(scope="module")
def ref():
    _ref_shim.install()
    tb.install_extension()                      # before the reference is imported
    with contextlib.redirect_stdout(io.StringIO()):
        import Models.SS2D.csms6s as cs
        import Models.vmamba as vm
        import Models.freq_mamba as fm
        import Trambav6
        from utils.loss import iou_loss
    return dict(cs=cs, vm=vm, fm=fm, Trambav6=Trambav6, iou_loss=iou_loss)


@contextlib.contextmanager
def patched(pairs):
    """setattr(obj, name, value) for the scope, then restore."""
    saved = [(o, n, getattr(o, n)) for o, n, _ in pairs]
    for o, n, v in pairs:
        setattr(o, n, v)
    try:
        yield
    finally:
        for o, n, v in saved:
            setattr(o, n, v)


def _run_and_check(ref, expect_calls):
    with installed() as lib:
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref["Trambav6"].bulid_model(deep_supervision=True, use_pretrain=False, img_size=384,
                                               dims=int(GOLD["dims"])).eval()
        det_fill(model)
        x = det_input((1, 3, 384, 384), 7)
        label = (det_input((1, 1, 384, 384), 8) > 0).float()
        outs = model(x)
        ups = [o if o.shape[-1] == 384 else F.interpolate(o, (384, 384), mode="bilinear") for o in outs]
        loss = sum(F.binary_cross_entropy_with_logits(o, label) + ref["iou_loss"](o, label) for o in ups)
        loss.backward()
    for name, n in expect_calls.items():
        assert lib.calls.count(name) == n, (name, lib.calls.count(name), n)
    rtol = 1e-3
    for i, o in enumerate(outs):
        a = o.detach().numpy()
        a = a if a.shape[-1] < 384 else a[..., ::4, ::4]
        g = GOLD[f"out{i}"]
        np.testing.assert_allclose(a, g, rtol=rtol, atol=rtol * np.abs(g).max(), err_msg=f"out{i}")
    assert abs(loss.item() - float(GOLD["loss"])) <= rtol * abs(float(GOLD["loss"]))
    params = dict(model.named_parameters())
    for k in GOLD.files:
        if k.startswith("grad:"):
            g = GOLD[k]
            np.testing.assert_allclose(params[k[5:]].grad.numpy(), g, rtol=2 * rtol,
                                       atol=2 * rtol * np.abs(g).max() + 1e-9, err_msg=k)
    return model


def test_level1_extension_module(ref):
    """Zero edits: the reference's own SelectiveScanOflex calls `selective_scan_cuda_oflex.fwd/.bwd`
    (csms6s.py:910,920-922) -- our module under that name."""
    cs = ref["cs"]
    assert sys.modules["selective_scan_cuda_oflex"] is tb.selective_scan_cuda_oflex
    with patched([(cs, "selective_scan_cuda_oflex", tb.selective_scan_cuda_oflex)]):   # the name csms6s.py:742 binds
        _run_and_check(ref, {"tramba_sscan_fwd": 33, "tramba_sscan_bwd": 33, "tramba_ss2d_fused_fwd": 0})


def test_level2_operator_classes(ref):
    """The autograd Functions swapped by name.  Defaults bound at def time (the encoder's
    scan=CrossScan) keep the reference's pure-torch classes; every explicit use goes through ours."""
    pairs = [(m, n, getattr(tb, n)) for n in OPS for m in (ref["cs"], ref["vm"], ref["fm"]) if hasattr(m, n)]
    with patched(pairs):
        _run_and_check(ref, {"tramba_sscan_fwd": 33, "tramba_sscan_bwd": 33})


def test_level3_modules_fused_core(ref):
    """SS2D and DCT2D replaced as modules: same constructor arguments, same parameter names, and
    the reference's state_dict keys (det_fill addresses parameters by key).  The reference passes its own
    scan / merge classes to the constructor; tb.SS2D adopts ours of the same name, so all 33 calls take the fused
    core exactly as in the harness."""
    vm, fm = ref["vm"], ref["fm"]
    ref_keys = None
    with contextlib.redirect_stdout(io.StringIO()):
        ref_keys = set(ref["Trambav6"].bulid_model(deep_supervision=True, use_pretrain=False, img_size=384,
                                                  dims=int(GOLD["dims"])).state_dict())
    with patched([(vm, "SS2D", tb.SS2D), (fm, "SS2D", tb.SS2D), (fm, "DCT2D", tb.DCT2D)]):
        model = _run_and_check(ref, {"tramba_ss2d_fused_fwd": 33, "tramba_ss2d_fused_bwd": 33,
                                     "tramba_sscan_fwd": 0, "tramba_dct2d_fwd": 3})
        assert set(model.state_dict()) == ref_keys
        ss2d = [m for m in model.modules() if isinstance(m, tb.SS2D)]
        assert len(ss2d) == 33
        assert all(m.scan.__module__ == "tramba_b200.ops" and m.merge.__module__ == "tramba_b200.ops" for m in ss2d)


def test_reference_checkpoint_loads_into_the_harness_and_the_dropin_modules(ref):
    """A state_dict saved from the REAL reference model (train.py:251-263) loads with strict=True into (a) the
    benchmark harness and (b) the reference model built on tb.SS2D / tb.DCT2D, value for value."""
    from tramba_b200.model import build_model
    dims = int(GOLD["dims"])
    with contextlib.redirect_stdout(io.StringIO()):
        src = ref["Trambav6"].bulid_model(deep_supervision=True, use_pretrain=False, img_size=384, dims=dims)
    det_fill(src)
    sd = src.state_dict()
    harness = build_model(img_size=384, dims=dims)
    res = harness.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    vm, fm = ref["vm"], ref["fm"]
    with patched([(vm, "SS2D", tb.SS2D), (fm, "SS2D", tb.SS2D), (fm, "DCT2D", tb.DCT2D)]):
        with contextlib.redirect_stdout(io.StringIO()):
            dropin = ref["Trambav6"].bulid_model(deep_supervision=True, use_pretrain=False, img_size=384, dims=dims)
    res = dropin.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in sd.items():
        assert torch.equal(harness.state_dict()[k].view_as(v), v), k
        assert torch.equal(dropin.state_dict()[k].view_as(v), v), k
