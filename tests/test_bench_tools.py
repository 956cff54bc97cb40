"""Host-side checks of the measurement plumbing (no GPU): kernel-name -> kind mapping used for the roofline,
and that profiles/roofline_traffic.json is what tools/make_traffic_json.py derives from the committed ncu list."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_kernel_name_to_kind():
    b = _load(os.path.join(ROOT, "bench.py"), "_bench")
    assert b.kind_of("tramba::ss2d_bwd_kernel<__nv_bfloat16, true>") == "ss2d_fused_bwd"
    assert b.kind_of("tramba::ss2d_fwd_kernel<float, false>") == "ss2d_fused_fwd"
    assert b.kind_of("tramba::dwconv_bwd_kernel<__nv_bfloat16, 7, 0>") == "dwconv7_bwd"
    assert b.kind_of("tramba::dwconv_fwd_kernel<float, 3, 1>") == "dwconv3_fwd"
    assert b.kind_of("tramba::ln2d_bwd_reg_kernel<float, __nv_bfloat16, 1>") == "layernorm2d_bwd"
    assert b.kind_of("at::native::vectorized_elementwise_kernel<4, ...>") is None
    assert b.kind_of("tramba::ss2d_bwd_kernel<__nv_bfloat16, true, 512, true>") == "ss2d_fused_bwd"
    assert b.kind_of("tramba::ss2d_fwd_kernel<float, true, 32, 1>") == "ss2d_fused_fwd"
    assert b.kind_of("tramba::batch_sum_kernel<__nv_bfloat16>") == "batch_sum"


def test_roofline_traffic_json_matches_committed_ncu_list(tmp_path):
    src = os.path.join(ROOT, "profiles", "r02_tramba_dram_per_launch.csv")
    dst = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not (os.path.exists(src) and os.path.exists(dst)):
        import pytest
        pytest.skip("profiles not present")
    m = _load(os.path.join(ROOT, "tools", "make_traffic_json.py"), "_mtj")
    out = tmp_path / "t.json"
    m.main(src, str(out))
    got, want = json.load(open(out)), json.load(open(dst))
    assert got["_launches"]["ss2d_fused_bwd"] == 33          # one step of Tramba-V: 33 SS2D calls
    assert abs(got["ss2d_fused_bwd"] - want["ss2d_fused_bwd"]) < 1.0
