"""ctypes binding of libpiccolo_b200.so (C ABI declared in include/piccolo_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is raised.
Build it with `python -c "import __graft_entry__ as g; g.build()"` or `make -C piccolo_b200/csrc`.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCL_LIB", os.path.join(_HERE, "libpiccolo_b200.so"))   # PCL_LIB: A/B testing of builds

c_float_p = ctypes.POINTER(ctypes.c_float)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); mirrors include/piccolo_b200.h one to one
SIGNATURES = {
    "pcl_abi_version": (ctypes.c_int, []),
    "pcl_last_error": (ctypes.c_char_p, []),
    "pcl_launch_count": (ctypes.c_int64, []),
    "pcl_cloud_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, c_void_pp]),
    "pcl_cloud_size": (ctypes.c_int64, [ctypes.c_void_p]),
    "pcl_cloud_bounds": (ctypes.c_int, [ctypes.c_void_p, c_float_p]),
    "pcl_cloud_destroy": (None, [ctypes.c_void_p]),
    "pcl_image_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, c_void_pp]),
    "pcl_image_format": (ctypes.c_int, [ctypes.c_void_p]),
    "pcl_image_destroy": (None, [ctypes.c_void_p]),
    "pcl_score": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_color_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_color_apply": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_color_mod_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_color_mod_apply": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_score_grid": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_topk": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_hist_rerank": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_hist_rerank_blocks": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_hist_rerank_finish": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_loss_fwd_bwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_refine_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, c_void_pp]),
    "pcl_refine_reset": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_refine_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "pcl_refine_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_refine_run_sharded": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_refine_debug_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "pcl_refine_debug_timeline": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "pcl_refine_destroy": (None, [ctypes.c_void_p]),
    "pcl_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "pcl_comm_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, c_void_pp]),
    "pcl_comm_handle": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_comm_connect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_comm_connect_ptrs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_comm_rank": (ctypes.c_int, [ctypes.c_void_p]),
    "pcl_comm_size": (ctypes.c_int, [ctypes.c_void_p]),
    "pcl_comm_barrier": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_comm_allgather_f32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pcl_comm_destroy": (None, [ctypes.c_void_p]),
}

_lib = None


class PiccoloError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the library once; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PiccoloError(
            f"{LIB_PATH} not found: the CUDA library is not built. Run `make -C piccolo_b200/csrc` "
            "(or __graft_entry__.build()). piccolo_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if "PCL_LIB" in os.environ and not hasattr(lib, name):
            continue                     # A/B testing of an older build (scripts/ab_refine.py): newer entry points are absent
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.pcl_abi_version() != 2 and "PCL_LIB" not in os.environ:
        raise PiccoloError("libpiccolo_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().pcl_last_error()
        raise PiccoloError(f"piccolo_b200 call failed ({rc}): {msg.decode() if msg else '?'}")


def set_option(name: str, value: int) -> None:
    """Tuning knob of the library (include/piccolo_b200.h: pcl_set_option); value < 0 restores the default."""
    check(load().pcl_set_option(name.encode(), int(value)))


def launch_count() -> int:
    return int(load().pcl_launch_count())
