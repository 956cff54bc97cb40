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
ORDER_STD, ORDER_DILATION, ORDER_WINDOW, ORDER_LINE, ORDER_CUSTOM = 0, 1, 2, 3, 4

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
    "tramba_dct_basis": (None, [_i, _fp]),
    "tramba_dct2d_fwd": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _vp]),
    "tramba_dct2d_bwd": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        from . import build as _build  # raises if nvcc is unavailable: no silent fallback
        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise RuntimeError(f"libtramba_b200: {lib.tramba_error_string(rc).decode()} (code {rc})")


def launch_count():
    return int(lib.tramba_launch_count())
