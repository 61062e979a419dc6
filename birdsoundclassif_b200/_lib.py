"""ctypes binding of libnbm_b200.so (include/nbm_b200.h).  There is NO fallback: if the
library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NBM_B200_LIB") or os.path.join(HERE, "libnbm_b200.so")   # override: diagnostic builds only


class NbmError(RuntimeError):
    pass


class FrontendParams(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("n_fft", C.c_int32), ("hop", C.c_int32), ("low_idx", C.c_int32),
                ("n_bins", C.c_int32), ("w_pix", C.c_int32), ("hop_spectro", C.c_int32), ("pad_mode", C.c_int32),
                ("stft_chunk", C.c_int64), ("min_level", C.c_double)]


class ProposalParams(C.Structure):
    _fields_ = [("A", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("img_width", C.c_float),
                ("img_height", C.c_float), ("min_size", C.c_float), ("nms_thresh", C.c_float),
                ("pre_nms_topN", C.c_int32), ("post_nms_topN", C.c_int32), ("rcnn_batch_size", C.c_int32)]


_p, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/nbm_b200.h declares
SIGNATURES = {
    "nbm_version": (C.c_int, []),
    "nbm_last_error": (C.c_char_p, []),
    "nbm_frontend_plan_create": (C.c_int, [C.POINTER(FrontendParams), C.POINTER(_p)]),
    "nbm_frontend_plan_destroy": (C.c_int, [_p]),
    "nbm_frontend_impl": (C.c_int, [_p]),
    "nbm_frontend_query": (C.c_int, [_p, _i64, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_sz)]),
    "nbm_frontend_query_batch": (C.c_int, [_p, C.POINTER(_i64), _i32, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_sz)]),
    "nbm_frontend_run": (C.c_int, [_p, _p, _i32, _i32, _i64, _p, _p, _p, _sz, _p]),
    "nbm_frontend_run_batch": (C.c_int, [_p, _p, _i32, _i32, C.POINTER(_i64), _i32, _p, _p, _p, _sz, _p]),
    "nbm_frontend_spectrogram_view": (C.c_int, [_p, C.POINTER(_i64), _i32, _i32, C.POINTER(_sz), C.POINTER(_i64)]),
    "nbm_frontend_set_profiling": (C.c_int, [_p, _i32]),
    "nbm_frontend_get_profile": (C.c_int, [_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64)]),
    "nbm_frontend_get_profile_kernels": (C.c_int, [_p, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "nbm_frontend_last_listed": (C.c_int, [_p, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "nbm_tiles_to_u8": (C.c_int, [_p, _i64, _p, _p]),
    "nbm_make_anchors": (C.c_int, [_i32, C.POINTER(C.c_double), _i32, C.POINTER(_i64), _i32, _i32, _i32, _i32, _p]),
    "nbm_decode_boxes": (C.c_int, [_p, _p, _i32, _i32, _i32, _f32, _f32, _f32, _p, _p, _p]),
    "nbm_nms_workspace_bytes": (_sz, [_i32, _i32]),
    "nbm_nms_greedy": (C.c_int, [_p, _p, _i32, _i32, _f32, _p, _p, _p, _sz, _p]),
    "nbm_proposals_workspace_bytes": (_sz, [C.POINTER(ProposalParams), _i32]),
    "nbm_proposals": (C.c_int, [C.POINTER(ProposalParams), _p, _p, _p, _i32, _p, _p, C.POINTER(_i32), _p, _sz, _p]),
    "nbm_proposals_async": (C.c_int, [C.POINTER(ProposalParams), _p, _p, _p, _i32, _p, _p, _p, _p, _sz, _p]),
    "nbm_final_detections": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _f32, _f32, _f32, _f32, _p, _p, _p, _p, _p]),
    "nbm_roi_pool": (C.c_int, [_p, _i32, _i32, C.POINTER(_p), C.POINTER(_i32), C.POINTER(_i32), _i32, _i32, _i32, _i32,
                               _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "nbm_merge_workspace_bytes": (_sz, [_i32]),
    "nbm_merge_detections": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _i64, _f32, _p, _p, _p, _p, _p, _sz, _p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library (once).  Raises NbmError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NbmError(f"{LIB_PATH} is missing: build it with `python -m birdsoundclassif_b200.build` "
                           "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise NbmError(f"{what} failed ({rc}): {lib().nbm_last_error().decode(errors='replace')}")
