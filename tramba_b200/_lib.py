"""ctypes binding of libtramba_b200.so (the C-ABI in include/tramba_b200.h).

There is no fallback: if the shared library is missing it is built with nvcc; if that is not
possible the import fails loudly.  Compute entry points need an sm_100 device and raise
RuntimeError (from the library's error code) otherwise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtramba_b200.so")

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_SILU = 0, 1, 2
ORDER_STD, ORDER_DILATION, ORDER_WINDOW, ORDER_LINE, ORDER_CUSTOM = 0, 1, 2, 3, 4
# ablation orders of the reference (csms6s.py:219-737); table-driven kernels only, never the fused core
ORDER_LINE4, ORDER_SPIRAL, ORDER_HILBERT, ORDER_DIAGONAL, ORDER_DS, ORDER_AB2, ORDER_AB1 = 5, 6, 7, 8, 9, 10, 11

_vp, _i, _i32p, _fp = ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/tramba_b200.h one to one
SIGNATURES = {
    "tramba_abi_version": (_i, []),
    "tramba_error_string": (ctypes.c_char_p, [_i]),
    "tramba_launch_count": (ctypes.c_uint64, []),
    "tramba_order_num_directions": (_i, [_i]),
    "tramba_order_table": (_i, [_i, _i, _i, _i, _i32p]),
    "tramba_order_plan_elems": (ctypes.c_size_t, [_i, _i]),
    "tramba_order_plan": (_i, [_i, _i, _i, _i, _i32p]),
    "tramba_order_plan_from_table": (_i, [_i, _i, _i32p, _i32p]),
    "tramba_cross_scan": (_i, [_vp, _vp, _i32p, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "tramba_cross_merge": (_i, [_vp, _vp, _i32p, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "tramba_sscan_ckpt_elems": (ctypes.c_size_t, [_i, _i, _i, _i, _i]),
    "tramba_sscan_fwd": (_i, [_vp, _vp, _fp, _vp, _vp, _fp, _fp, _vp, _fp] + [_i] * 8 + [_vp]),
    "tramba_sscan_bwd": (_i, [_vp, _vp, _fp, _vp, _vp, _fp, _fp, _vp, _fp, _vp, _vp, _fp, _fp, _fp, _fp, _fp]
                         + [_i] * 8 + [_vp]),
    "tramba_order_table16": (_i, [_i, _i, _i, _i, _vp]),
    "tramba_order_table16_skewed": (_i, [_i, _i, _i, _i, _vp]),
    "tramba_ss2d_fused_supported": (_i, [_i] * 8),
    "tramba_ss2d_fused_fwd": (_i, [_vp, _vp, _vp, _vp, _fp, _fp, _fp, _vp, _fp] + [_i] * 9 + [_vp]),
    "tramba_order_sched_len": (ctypes.c_size_t, [_i, _i, _i, _i]),
    "tramba_order_sched": (_i, [_i, _i, _i, _i, _vp, _i32p]),
    "tramba_ss2d_fused_bwd": (_i, [_vp, _vp, _vp, _vp, _fp, _fp, _fp, _vp, _i, _i32p, _fp, _fp, _vp, _fp, _fp, _fp,
                                      _fp, _fp] + [_i] * 9 + [_vp]),
    "tramba_layernorm2d_tiles": (_i, [_i, _i, _i]),
    "tramba_layernorm2d_fwd": (_i, [_vp, _fp, _fp, _vp, _fp, _fp, _i, _i, _i, ctypes.c_float, _i, _i, _i, _vp]),
    "tramba_layernorm2d_bwd": (_i, [_vp, _vp, _fp, _fp, _fp, _fp, _vp, _fp, _fp] + [_i] * 6 + [_vp]),
    "tramba_proj_gather": (_i, [_vp, _vp, _vp, _vp, _i32p] + [_i] * 6 + [_vp]),
    "tramba_proj_scatter": (_i, [_vp, _vp, _vp, _vp, _i32p] + [_i] * 7 + [_vp]),
    "tramba_batch_sum": (_i, [_vp, _vp, _i, _i, ctypes.c_longlong, _i, _vp]),
    "tramba_dwconv2d_supported": (_i, [_i] * 5),
    "tramba_dwconv2d_fwd": (_i, [_vp, _fp, _fp, _vp] + [_i] * 7 + [_vp]),
    "tramba_dwconv2d_bwd": (_i, [_vp, _fp, _fp, _vp, _vp, _fp] + [_i] * 7 + [_vp]),
    "tramba_dct_basis": (None, [_i, _fp]),
    "tramba_dct2d_supported": (_i, [_i, _i]),
    "tramba_dct2d_fwd": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _vp]),
    "tramba_dct2d_bwd": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _vp]),
}


def _load():
    import importlib.util
    # build.py is loaded by path so that building never needs the (possibly stale) library itself
    spec = importlib.util.spec_from_file_location("_tramba_b200_build", os.path.join(_HERE, "build.py"))
    _build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(_build)
    # one builder at a time: under torchrun every rank imports the package at the same moment
    import fcntl

    def ensure_built():
        if not os.path.exists(LIB_PATH) or (_build.sources_newer_than_lib() and os.path.exists(_build.NVCC)):
            _build.build()  # raises if nvcc is unavailable or fails: no silent fallback

    try:
        lock = open(os.path.join(_HERE, ".build.lock"), "w")
    except OSError:          # read-only tree: nothing to serialise, the library must already be there
        lock = None
    try:
        if lock is not None:
            fcntl.flock(lock, fcntl.LOCK_EX)
        ensure_built()
    finally:
        if lock is not None:
            fcntl.flock(lock, fcntl.LOCK_UN)
            lock.close()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


class _NotLoaded:
    """TRAMBA_B200_NO_LIB=1 (set by `bench.py --impl reference` only): the process must not map the CUDA library --
    the reference arm times the CPU path and its record must show no native code of ours.  Any use raises."""

    def __getattr__(self, name):
        raise RuntimeError(f"tramba_b200: libtramba_b200.so deliberately not loaded (TRAMBA_B200_NO_LIB=1); "
                           f"`{name}` is not available in this process")


lib = _NotLoaded() if os.environ.get("TRAMBA_B200_NO_LIB", "") == "1" else _load()


def check(rc):
    if rc != 0:
        raise RuntimeError(f"libtramba_b200: {lib.tramba_error_string(rc).decode()} (code {rc})")


def launch_count():
    return int(lib.tramba_launch_count())


# ---------------------------------------------------------------------------------------
# Per-launch device timing (bench.py's roofline leg).  When `prof` is a list, every C-ABI compute
# call made through the Python host is bracketed by CUDA events recorded on the launching stream
# and appended as (kind, algorithmic_bytes, start_event, end_event); bench.py reads the elapsed
# times after its own synchronize.  Off (None) by default: zero overhead on the product path.
# ---------------------------------------------------------------------------------------
prof = None
# TRAMBA_NVTX=1: every C-ABI compute call is wrapped in an NVTX range named after its kind ("ss2d_fused_bwd", ...)
# so that `ncu --nvtx --nvtx-include "ss2d_fused_bwd/"` (or nsys) can select call sites by stage.  Off by default.
NVTX = os.environ.get("TRAMBA_NVTX", "") == "1"


class timed:
    __slots__ = ("kind", "nbytes", "ev")

    def __init__(self, kind, nbytes):
        self.kind, self.nbytes, self.ev = kind, nbytes, None

    def __enter__(self):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_push(self.kind)
        if prof is not None:
            import torch
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            self.ev[1].record()
            prof.append((self.kind, self.nbytes, self.ev[0], self.ev[1]))
        if NVTX:
            import torch
            torch.cuda.nvtx.range_pop()
        return False
