"""ctypes binding of libdrt_b200.so (include/drt_b200.h).  There is NO fallback: if the CUDA
library is missing or fails, every entry point raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DRT_B200_LIB: developer override used to A/B kernel variants built with different -D flags
SO_PATH = os.environ.get("DRT_B200_LIB") or os.path.join(_HERE, "_C", "libdrt_b200.so")

_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); must list every symbol include/drt_b200.h declares
SIGNATURES = {
    "drt_version": (C.c_int, []),
    "drt_last_error": (C.c_char_p, []),
    "drt_kernel_launches": (C.c_uint64, []),
    "drt_tuning_set": (C.c_int, [C.c_char_p, C.c_longlong]),
    "drt_tuning_get": (C.c_longlong, [C.c_char_p]),
    "drt_bvh_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "drt_bvh_destroy": (C.c_int, [_vp]),
    "drt_bvh_build": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _vp]),
    "drt_bvh_build_f64": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _vp]),
    "drt_bvh_update_vert": (C.c_int, [_vp, _vp, _vp, _i32, C.c_int, _vp]),
    "drt_bvh_bad_indices": (C.c_int, [_vp, _vp, C.POINTER(C.c_int)]),
    "drt_bvh_info": (C.c_int, [_vp, C.POINTER(_i64)]),
    "drt_bvh_last_counts": (C.c_int, [_vp, _vp, C.POINTER(_i64)]),
    "drt_bvh_set_image_size": (C.c_int, [_vp, _i32, _i32]),
    "drt_closest_hit": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _vp]),
    "drt_trace_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drt_trace_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drt_trace_fwd_smooth": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drt_trace_bwd_smooth": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drt_plane_hit": (C.c_int, [_vp, _vp, _vp, _i64, C.POINTER(_f64), _vp, _vp, _vp]),
    "drt_plane_hit_bwd": (C.c_int, [_vp, _vp, _vp, _i64, C.POINTER(_f64), _vp, _vp, _vp, _vp]),
    "drt_ray_loss_grad": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "drt_ray_loss_grad_rec": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "drt_ray_loss_step": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _f64, _f64, C.c_int, _vp, _vp, _vp, _vp, _i64, _i32, _i32,
                                    _vp, _vp, _vp, _vp, _vp]),
    "drt_ray_loss_step_beams": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _f64, _f64, C.c_int, _vp, _vp, _vp, _vp, _i64, _i32, _i32,
                                          _vp, _vp, _vp, _vp, _vp, _vp]),
    "drt_tile_beams_floats": (_i64, [_i64]),
    "drt_tile_beams": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp]),
    "drt_generate_rays": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "drt_silhouette_classify": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "drt_silhouette_sample": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "drt_silhouette_backward": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _i64, _vp, _vp]),
    "drt_silhouette_loss": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, C.c_int, _vp, _vp, _vp, _vp]),
    "drt_dihedral_loss": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "drt_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, _i64, C.POINTER(_vp)]),
    "drt_comm_handle": (C.c_int, [_vp, C.c_char_p]),
    "drt_comm_connect": (C.c_int, [_vp, C.c_char_p]),
    "drt_comm_allreduce_sum_f64": (C.c_int, [_vp, _vp, _i64, _vp]),
    "drt_comm_status": (C.c_int, [_vp, _vp, C.POINTER(C.c_int)]),
    "drt_comm_destroy": (C.c_int, [_vp]),
}

_lib = None


class DrtError(RuntimeError):
    pass


def load():
    """Loads the in-tree CUDA library.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise DrtError(
                f"{SO_PATH} not found: build it with `python -m drt_b200.build` (nvcc, sm_100a). "
                "drt_b200 has no CPU or PyTorch fallback path.")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise DrtError(f"libdrt_b200 error {rc}: {load().drt_last_error().decode()}")


def call(name, *args):
    check(getattr(load(), name)(*args))
