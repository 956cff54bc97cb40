"""The candidate build (libtramba_b200_next.so = same sources + csrc/next/ with -DTRAMBA_NEXT) keeps the product's
ABI, really contains the compile-time-geometry kernels, and the PRODUCT library does not contain them (the default
path is the GPU-verified one).  CPU only: builds with nvcc, inspects symbols; no compute."""
import importlib.util
import os
import subprocess

import pytest

from conftest import ROOT
from tramba_b200 import _lib


@pytest.fixture(scope="module")
def next_lib():
    spec = importlib.util.spec_from_file_location("_tb_build", os.path.join(ROOT, "tramba_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(next_variant=True)          # incremental: seconds when the objects are fresh


def _exports(path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    return sorted(l.split()[-1] for l in out.splitlines() if " T " in l)


def _kernel_names(path):
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    out = subprocess.run(["cuobjdump", "-symbols", path], capture_output=True, text=True).stdout
    return out


def test_candidate_library_has_the_same_abi(next_lib):
    base = os.path.join(os.path.dirname(next_lib), "libtramba_b200.so")
    assert _exports(next_lib) == _exports(base) == sorted(_lib.SIGNATURES)


def test_candidate_kernels_only_in_the_candidate_library(next_lib):
    base = os.path.join(os.path.dirname(next_lib), "libtramba_b200.so")
    cand_syms, base_syms = _kernel_names(next_lib), _kernel_names(base)
    for k in ("ss2d_fwd_cg_kernel", "ss2d_bwd_cg_kernel", "merge_skewed_cg_kernel"):
        assert k in cand_syms, k
        assert k not in base_syms, k             # the default build is the verified kernels only
    for k in ("ss2d_fwd_kernel", "ss2d_bwd_kernel"):   # the generic kernels stay in both (fallback for other shapes)
        assert k in cand_syms and k in base_syms


def test_default_import_loads_the_product_library():
    assert os.environ.get("TRAMBA_B200_LIB", "") == ""
    assert os.path.basename(_lib.LIB_PATH) == "libtramba_b200.so"
