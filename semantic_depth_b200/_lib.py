"""ctypes binding of libsd_fusion.so -- the C ABI declared in include/sd_fusion.h.

There is no CPU fallback: if the library is missing or cannot be loaded this module raises, and
every entry point of the package fails with it (the product path is the CUDA path or nothing).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsd_fusion.so")

SD_OK = 0
SD_NUM_COUNTS = 17
COUNT_NAMES = (
    "road_gather", "fence_gather", "road_z", "road_mad_y", "road_mad_x", "road_plane", "road_sor", "road_ror",
    "road_slab", "fence_mad_y", "fence_abs_z", "left_split", "right_split", "left_mad_x", "left_plane",
    "right_mad_x", "right_plane",
)
assert len(COUNT_NAMES) == SD_NUM_COUNTS

STAGE_NAMES = ("pixel", "road_mad", "road_plane", "road_grid", "road_knn", "road_ror", "rw", "fences", "answers")

(PRED_LT, PRED_ABS_LT, PRED_MAD, PRED_PLANE, PRED_GT, PRED_SLAB, PRED_SOR, PRED_ROR) = range(8)


class SdCamera(C.Structure):
    _fields_ = [("q03", C.c_float), ("q13", C.c_float), ("q23", C.c_float), ("q32", C.c_float),
                ("disparity_mult", C.c_float)]


class SdParams(C.Structure):
    _fields_ = [
        ("prob_thr", C.c_double),
        ("road_z_to_meter", C.c_float), ("road_mad_y_thr", C.c_float), ("road_mad_x_thr", C.c_float),
        ("fence_mad_y_thr", C.c_float), ("fence_abs_z_thr", C.c_float), ("left_mad_x_thr", C.c_float),
        ("right_mad_x_thr", C.c_float),
        ("sor_nb_neighbors", C.c_int32),
        ("road_plane_thr", C.c_double), ("fence_plane_thr", C.c_double), ("sor_std_ratio", C.c_double),
        ("ror_radius", C.c_double), ("slab_lo", C.c_double), ("slab_hi", C.c_double), ("depth", C.c_double),
        ("ror_nb_points", C.c_int32), ("use_sor", C.c_int32), ("use_ror", C.c_int32), ("approach_both", C.c_int32),
        ("label_mode", C.c_int32), ("pad_", C.c_int32),
    ]


class SdFrameResult(C.Structure):
    _fields_ = [
        ("rw", C.c_double), ("f2f", C.c_double), ("xl", C.c_double), ("xr", C.c_double),
        ("left_pt", C.c_double * 3), ("right_pt", C.c_double * 3),
        ("road_coeff", C.c_double * 4), ("left_coeff", C.c_double * 4), ("right_coeff", C.c_double * 4),
        ("sor_mean", C.c_double), ("sor_std", C.c_double), ("sor_thr", C.c_double),
        ("median", C.c_float * 5), ("mad", C.c_float * 5), ("fence_mean_x", C.c_float),
        ("status", C.c_uint32),
        ("counts", C.c_int32 * SD_NUM_COUNTS),
        ("ransac_best", C.c_int32 * 3),
    ]


class SdFcnHeadWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("conv3_w", "conv3_b", "conv4_w", "conv4_b", "conv7_w", "conv7_b",
                                         "deconv1_w", "deconv1_b", "deconv2_w", "deconv2_b")]


class SdPredicate(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("axis", C.c_int32), ("ia", C.c_int32), ("use_f32", C.c_int32),
        ("fa", C.c_float), ("f0", C.c_float), ("f1", C.c_float), ("pad_", C.c_float),
        ("da", C.c_double), ("d0", C.c_double), ("d1", C.c_double), ("d2", C.c_double),
        ("d_aux", C.c_void_p),
    ]


_P = C.c_void_p
_I = C.c_int
# name -> (restype, argtypes): exactly the prototypes of include/sd_fusion.h
SIGNATURES = {
    "sd_abi_version": (C.c_int, []),
    "sd_last_error": (C.c_char_p, []),
    "sd_default_params": (None, [C.POINTER(SdParams), C.c_double]),
    "sd_ws_bytes": (C.c_size_t, [_I, _I, _I, _I]),
    "sd_ws_create": (_I, [C.POINTER(_P), _P, C.c_size_t, _I, _I, _I, _I, _P]),
    "sd_ws_destroy": (None, [_P]),
    "sd_pixel_fuse": (_I, [_P, _P, _P, _P, _I, _I, _I, C.POINTER(SdCamera), C.c_double, C.c_float, _I,
                           _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "sd_pixel_fuse_scores": (_I, [_P, _P, _P, _P, _I, _I, _I, C.POINTER(SdCamera), C.c_double, C.c_float, _I,
                                  _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "sd_fuse_frames_scores": (_I, [_P, _P, _P, _P, _I, _I, _I, C.POINTER(SdCamera), C.POINTER(SdParams), _P, _P, _P]),
    "sd_ply_rows": (_I, [_P, _P, _P, _P, _I, _P, C.c_ulonglong, C.POINTER(C.c_ulonglong), _P, _P]),
    "sd_ply_rows_f64": (_I, [_P, _P, _P, _P, _I, _P, C.c_ulonglong, C.POINTER(C.c_ulonglong), _P, _P]),
    "sd_resize_cubic_u8": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _P]),
    "sd_overlay_masks": (_I, [_P, _P, _I, _I, _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P, _P, _P]),
    "sd_draw_banner": (_I, [_P, _I, _I, _I, _P, _I, _P, _I, _I, _I, _P, _I, _P]),
    "sd_median_mad": (_I, [_P, _I, C.POINTER(C.c_float), _P, _P]),
    "sd_filter": (_I, [_P, _P, _P, _P, _I, C.POINTER(SdPredicate), _P, _P, _P, _P, C.POINTER(C.c_int32), _P, _P]),
    "sd_plane_fit": (_I, [_P, _P, _P, _I, _I, C.POINTER(C.c_double), C.POINTER(C.c_int32), _P, _P]),
    "sd_mean_f32": (_I, [_P, _I, C.POINTER(C.c_float), _P, _P]),
    "sd_slab_minmax": (_I, [_P, _P, _I, C.c_double, C.c_double, _I, C.POINTER(C.c_float), C.POINTER(C.c_float),
                            C.POINTER(C.c_int32), _P, _P]),
    "sd_knn_mean_distance": (_I, [_P, _P, _P, _I, _I, C.c_double, _P, C.POINTER(C.c_double), _P, _P]),
    "sd_radius_count": (_I, [_P, _P, _P, _I, C.c_double, _I, _P, _P, _P]),
    "sd_ransac_score": (_I, [_P, _P, _P, _I, _I, C.c_double, _P, _I, _P, C.POINTER(C.c_int32),
                             C.POINTER(C.c_double), _P, _P]),
    "sd_fuse_frames": (_I, [_P, _P, _I, _I, _I, C.POINTER(SdCamera), C.POINTER(SdParams), _P, _P, _P, _I, _P, _P, _P]),
    "sd_fuse_kernel_count": (_I, [C.POINTER(SdParams), _I]),
    "sd_fuse_frames_host": (_I, [_P, _P, _I, _I, _I, C.POINTER(SdCamera), C.POINTER(SdParams), _P, _P, _P, _P, _P, _P]),
    "sd_ws_enable_timing": (_I, [_P, _I]),
    "sd_ws_set_stage_mask": (_I, [_P, _I]),
    "sd_ws_stage_elapsed_ms": (_I, [_P, _I, C.POINTER(C.c_float)]),
    "sd_ws_stage_times": (_I, [_P, C.POINTER(C.c_float)]),
    "sd_fcn8s_head_scratch_bytes": (C.c_size_t, [_I, _I, _I]),
    "sd_fcn8s_head": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, C.POINTER(SdFcnHeadWeights), _P, C.c_size_t, _P, _P]),
    "sd_ws_cloud": (_I, [_P, _I, _I, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "sd_ws_stage_src": (_I, [_P, _I, _I, C.POINTER(_P)]),
    "sd_ws_stage_alive": (_I, [_P, _I, _I, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
}

_lib = None


class SdError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libsd_fusion.so (built by semantic_depth_b200.build); raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SdError(f"{LIB_PATH} not found: build it with `python -m semantic_depth_b200.build` "
                      "(there is no CPU fallback for the fusion path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.sd_abi_version() != 2:
        raise SdError("libsd_fusion.so ABI version mismatch")
    _lib = lib
    return lib


OPS_PATH = os.path.join(os.path.dirname(LIB_PATH), "libsd_torch_ops.so")
_ops = None


def load_ops():
    """The TORCH_LIBRARY op layer over the C ABI (csrc/sd_torch_ops.cpp): ``torch.ops.sd_fusion``.  It validates device,
    dtype and contiguity of every tensor in C++ (RuntimeError) and forwards to sd_fuse_frames / sd_fuse_frames_scores."""
    global _ops
    if _ops is not None:
        return _ops
    import torch
    load()                                   # the C ABI library first: same checks, same error if it is missing
    if not os.path.exists(OPS_PATH):
        raise SdError(f"{OPS_PATH} not found: build it with `python -m semantic_depth_b200.build`")
    torch.ops.load_library(OPS_PATH)
    ops = torch.ops.sd_fusion
    if int(ops.abi_version()) != 2 or int(ops.result_bytes()) != C.sizeof(SdFrameResult):
        raise SdError("libsd_torch_ops.so does not match libsd_fusion.so / the ctypes structs")
    _ops = ops
    return ops


def struct_tensor(st):
    """A ctypes struct of the ABI as the CPU uint8 tensor the op layer takes."""
    import torch
    return torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8)


def check(rc: int, what: str = "") -> None:
    if rc != SD_OK:
        msg = load().sd_last_error()
        raise SdError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
