"""ctypes binding of the C ABI declared in include/ypb200.h (the only way Python reaches the kernels).

There is deliberately no fallback: if libypb200.so is missing or cannot be loaded, importing the
engine raises.
"""

import ctypes as C
import os

from . import build as _build

_LIB = None

c_void_p, c_int, c_float, c_size_t, c_char_p = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_char_p
c_int64, c_uint32, c_double = C.c_int64, C.c_uint32, C.c_double


class InferParams(C.Structure):
    _fields_ = [("conf", c_float), ("iou", c_float), ("max_det", c_int), ("agnostic_nms", c_int),
                ("class_mask", c_void_p)]


# name -> (restype, argtypes); mirrors include/ypb200.h one to one
SIGNATURES = {
    "ypb_version": (c_int, []),
    "ypb_last_error": (c_char_p, []),
    "ypb_engine_create": (c_int, [c_char_p, c_int, C.POINTER(c_void_p)]),
    "ypb_engine_destroy": (None, [c_void_p]),
    "ypb_weight_count": (c_int, [c_void_p]),
    "ypb_weight_info": (c_int, [c_void_p, c_int, C.POINTER(c_char_p), C.POINTER(c_int), C.POINTER(c_int64 * 4),
                                C.POINTER(c_int)]),
    "ypb_load_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "ypb_finalize_weights": (c_int, [c_void_p, c_int]),
    "ypb_plan": (c_int, [c_void_p, c_int, c_int, c_int, C.POINTER(c_size_t)]),
    "ypb_bind_workspace": (c_int, [c_void_p, c_void_p, c_size_t]),
    "ypb_num_anchors": (c_int, [c_void_p]),
    "ypb_num_classes": (c_int, [c_void_p]),
    "ypb_num_mask_coefs": (c_int, [c_void_p]),
    "ypb_kernel_launches": (c_int, [c_void_p]),
    "ypb_conv_flops": (c_double, [c_void_p]),
    "ypb_infer": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(InferParams), c_void_p, c_void_p,
                          c_void_p, c_void_p, c_void_p]),
    "ypb_infer_profile": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(InferParams), c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "ypb_op_count": (c_int, [c_void_p]),
    "ypb_op_info": (c_int, [c_void_p, c_int, C.POINTER(c_char_p), C.POINTER(c_int), C.POINTER(c_double),
                            C.POINTER(c_double)]),
    "ypb_masks": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_int, c_void_p]),
    "ypb_masks_ex": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int, c_void_p, c_void_p, c_void_p]),
    "ypb_proto_info": (c_int, [c_void_p, C.POINTER(c_size_t), C.POINTER(c_size_t)]),
    "ypb_select_info": (c_int, [c_void_p, C.POINTER(c_size_t), C.POINTER(c_int)]),
    "ypb_device_error": (c_int, [c_void_p, C.POINTER(c_uint32)]),
    "ypb_device_error_async": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ypb_view_count": (c_int, [c_void_p]),
    "ypb_view_info": (c_int, [c_void_p, c_int, C.POINTER(c_char_p), C.POINTER(c_size_t), C.POINTER(c_int),
                              C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int)]),
    "ypb_set_conv_impl": (c_int, [c_void_p, c_int]),
    "ypb_set_graph": (c_int, [c_void_p, c_int]),
    "ypb_conv2d_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int]),
    "ypb_nms": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_int, c_void_p,
                        c_void_p, c_void_p]),
    "ypb_nms_scratch_bytes": (c_size_t, [c_int, c_int]),
    "ypb_stage_frames": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "ypb_stage_frames_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]),
    "ypb_stage_frames_gated": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "ypb_mask_min_rect": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ypb_letterbox_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "ypb_index_masks": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ypb_index_masks_boxed": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p]),
    "ypb_index_masks_resized": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "ypb_jpeg_info": (c_int, [c_void_p, c_size_t, C.POINTER(c_int), C.POINTER(c_int)]),
    "ypb_jpeg_decode_bgr": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_int, c_int]),
    "ypb_mailbox_create": (c_int, [c_int, c_size_t, C.POINTER(c_void_p), c_void_p]),
    "ypb_mailbox_destroy": (c_int, [c_int, c_void_p]),
    "ypb_mailbox_open": (c_int, [c_int, c_void_p, C.POINTER(c_void_p)]),
    "ypb_mailbox_close": (c_int, [c_int, c_void_p]),
    "ypb_peer_copy": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "ypb_host_is_pinned": (c_int, [c_void_p, C.POINTER(c_int)]),
    "ypb_hosts_are_pinned": (c_int, [c_void_p, c_int, C.POINTER(c_int)]),
    "ypb_h2d_frames": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
    "ypb_is_diag_build": (c_int, []),
}

# include/ypb200_diag.h: exported by libypb200_diag.so only (debugging twins, micro-benchmarks)
DIAG_SIGNATURES = {
    "ypb_conv_bench": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                               c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                               C.POINTER(c_float), C.c_char_p, c_int]),
    "ypb_debug_prof": (c_int, [c_void_p, c_int]),
    "ypb_mma_bench": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, C.POINTER(c_float)]),
    "ypb_latency_probe": (c_int, [c_void_p]),
    "ypb_tma_bench": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, C.POINTER(c_float),
                              C.POINTER(c_double)]),
}


class YpbError(RuntimeError):
    pass


def _attach(handle, table):
    for name, (res, args) in table.items():
        fn = getattr(handle, name)
        fn.restype, fn.argtypes = res, args


def lib():
    """Load (building first if sources are newer) libypb200.so and attach prototypes."""
    global _LIB
    if _LIB is None:
        path = _build.ensure_built()
        if os.environ.get("YPB_LIB"):  # diagnostics: e.g. the profiling build libypb200_prof.so (tools/conv_layers.py)
            path = os.environ["YPB_LIB"]
        if not os.path.exists(path):
            raise YpbError(f"{path} missing: the CUDA extension is required, there is no fallback")
        handle = C.CDLL(path)
        _attach(handle, SIGNATURES)
        if handle.ypb_is_diag_build():
            _attach(handle, DIAG_SIGNATURES)
        _LIB = handle
    return _LIB


_DIAG = None


def diag_lib():
    """libypb200_diag.so: the product translation unit plus debugging twins and micro-benchmarks (built on demand,
    in-tree).  Tests of the twins and the tools/ scripts use it; the product path never does."""
    global _DIAG
    if _DIAG is None:
        if lib().ypb_is_diag_build():
            _DIAG = lib()
        else:
            handle = C.CDLL(_build.build(diag=True))
            _attach(handle, SIGNATURES)
            _attach(handle, DIAG_SIGNATURES)
            _DIAG = handle
    return _DIAG


def check(rc):
    if rc != 0:
        raise YpbError(f"ypb200 error {rc}: {lib().ypb_last_error().decode()}")
