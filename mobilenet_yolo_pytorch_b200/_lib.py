"""ctypes binding of libb200yolo.so (include/b200yolo.h).

The library is the product: there is no Python / PyTorch / CPU fallback.  If the
shared object is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200yolo.so")

_lock = threading.Lock()
_lib = None

c_f32p = C.c_void_p  # raw device / host addresses are passed as integers
c_i32p = C.c_void_p

# every exported symbol of include/b200yolo.h: (restype, argtypes)
SIGNATURES = {
    "b200yolo_version": (C.c_int, []),
    "b200yolo_last_error": (C.c_char_p, []),
    "b200yolo_launch_count": (C.c_ulonglong, []),
    "b200yolo_max_cells": (C.c_int, [C.c_int]),
    "b200yolo_debug_phase_stamps": (None, [C.c_void_p]),
    "b200yolo_debug_set_flags": (None, [C.c_int]),
    "b200yolo_decode_head": (C.c_int, [c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_f32p, C.c_float,
                                       c_f32p, c_i32p, c_i32p, C.c_void_p]),
    "b200yolo_nms": (C.c_int, [c_f32p, c_i32p, C.c_int, c_f32p, c_i32p, C.c_int, C.c_int, C.c_int, C.c_double,
                               c_f32p, c_i32p, c_i32p, C.c_void_p]),
    "b200yolo_decode_nms": (C.c_int, [c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      c_f32p, C.c_float, C.c_double, c_f32p, c_i32p, c_i32p, C.c_void_p]),
    "b200yolo_decode_nms_batches": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              c_f32p, C.c_float, C.c_double, C.c_void_p]),
    "b200yolo_plan_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_float, C.c_double, C.POINTER(C.c_void_p)]),
    "b200yolo_plan_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200yolo_plan_destroy": (C.c_int, [C.c_void_p]),
    "b200yolo_mc_supported": (C.c_int, [C.c_int]),
    "b200yolo_mc_create": (C.c_int, [C.c_size_t, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "b200yolo_mc_import": (C.c_int, [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "b200yolo_mc_add_device": (C.c_int, [C.c_void_p]),
    "b200yolo_mc_bind": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "b200yolo_mc_free": (C.c_int, [C.c_void_p]),
    "b200yolo_set_inputs_ready": (None, [C.c_int]),
    "b200yolo_set_exact_decode": (None, [C.c_int]),
    "b200yolo_decode_nms_nhwc": (C.c_int, [c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           c_f32p, C.c_float, C.c_double, c_f32p, c_i32p, c_i32p, C.c_void_p]),
    "b200yolo_decode_nms_large_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "b200yolo_decode_nms_large": (C.c_int, [c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            c_f32p, C.c_float, C.c_double, c_f32p, c_i32p, c_i32p, C.c_void_p, C.c_size_t,
                                            C.c_void_p]),
    "b200yolo_decode_nms_gather": (C.c_int, [c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             c_f32p, C.c_float, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                             C.c_void_p]),
    "b200yolo_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "b200yolo_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "b200yolo_peer_signal": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b200yolo_peer_wait": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "b200yolo_peer_fence": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "b200yolo_decode_nms_gather_steps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                   C.c_int, C.c_int, C.c_int, c_f32p, C.c_float, C.c_double, C.c_void_p]),
    "b200yolo_peer_close": (C.c_int, [C.c_void_p]),
    "b200yolo_peer_free": (C.c_int, [C.c_void_p]),
    "b200yolo_decode_nms_host": (C.c_int, [c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, c_f32p, C.c_float, C.c_double, c_f32p, c_i32p, C.c_int]),
    "b200yolo_decode_nms_host_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b200yolo_decode_nms_host_ws": (C.c_int, [c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, c_f32p, C.c_float, C.c_double, c_f32p, c_i32p, C.c_void_p, C.c_size_t, C.c_int]),
    "b200yolo_host_last_d2h_bytes": (C.c_size_t, []),
    "b200yolo_compact_rows": (C.c_int, [c_f32p, c_i32p, C.c_int, C.c_int, c_f32p, c_i32p, C.c_void_p]),
    "b200yolo_pairwise": (C.c_int, [c_f32p, C.c_int, c_f32p, C.c_int, C.c_int, c_f32p, C.c_void_p]),
    "b200yolo_target_loss_workspace_bytes": (C.c_size_t, [C.c_int]),
    "b200yolo_target_loss": (C.c_int, [c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_f32p, C.c_int, c_i32p,
                                       c_f32p, c_i32p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, c_i32p, c_f32p,
                                       c_i32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200yolo_target_loss_backward": (C.c_int, [c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_f32p, C.c_int, c_i32p,
                                                c_f32p, c_i32p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_float, c_f32p, c_f32p, C.c_void_p]),
    "b200yolo_loss_finalize": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p]),
    "b200yolo_loss_finalize_dev": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "b200yolo_seg_loss_workspace_bytes": (C.c_size_t, []),
    "b200yolo_seg_loss": (C.c_int, [c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.c_void_p]),
    "b200yolo_seg_loss_backward": (C.c_int, [c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, c_f32p, c_f32p, C.c_void_p]),
    "b200yolo_seg_sigmoid": (C.c_int, [c_f32p, C.c_longlong, c_f32p, C.c_void_p]),
    "b200yolo_map_eval_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "b200yolo_map_eval": (C.c_int, [c_f32p, c_i32p, c_f32p, c_i32p, C.c_int, c_f32p, c_i32p, C.c_void_p, c_i32p, C.c_int,
                                    C.c_int, C.c_int, C.c_float, c_f32p, C.c_int, c_f32p, c_f32p, c_f32p, C.c_void_p,
                                    C.c_size_t, C.c_void_p]),
}

# partial-sum slots (enum in include/b200yolo.h)
S_SQW, S_W, S_IOU_SQ, S_IOU_W, S_NASSIGN, S_OBJ, S_CONF_ALL, S_CLS, S_IOU, S_RECALL, S_NCELLS, S_NIMG = range(12)
S_COUNT = 16


class Batch(C.Structure):
    """struct b200yolo_batch (include/b200yolo.h)."""
    _fields_ = [("head0", C.c_void_p), ("head1", C.c_void_p), ("out", C.c_void_p), ("out_count", C.c_void_p),
                ("out_idx", C.c_void_p)]


class Gather(C.Structure):
    """struct b200yolo_gather (include/b200yolo.h)."""
    _fields_ = [("R", C.c_int), ("rank", C.c_int), ("peer_out", (C.c_void_p * 8) * 3), ("peer_count", (C.c_void_p * 8) * 3),
                ("peer_flags", C.c_void_p * 8), ("timed_out", C.c_void_p), ("timeout_s", C.c_double), ("multicast", C.c_int)]


class B200YoloError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200yolo error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    """Load the shared library (build it with `python -m mobilenet_yolo_pytorch_b200.build`
    or `__graft_entry__.build()`).  Raises if it is not there -- no fallback."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build the CUDA extension first "
                    "(python -m mobilenet_yolo_pytorch_b200.build); there is no CPU fallback")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the symbol is not exported
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise B200YoloError(rc, load().b200yolo_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().b200yolo_launch_count())
