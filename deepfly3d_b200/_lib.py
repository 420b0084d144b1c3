"""ctypes binding of libdf3d_b200.so (the C ABI declared in include/df3d_b200.h).

There is no fallback: if the shared library is missing the import fails loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdf3d_b200.so")

DF3D_MAX_CAMS = 8
DF3D_EUNSUPPORTED = -4


class BAOpts(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("ftol", C.c_double), ("xtol", C.c_double), ("gtol", C.c_double), ("solver", C.c_int)]


class BAReport(C.Structure):
    _fields_ = [
        ("cost0", C.c_double), ("cost", C.c_double), ("reg", C.c_double),
        ("iters", C.c_int32), ("accepted", C.c_int32), ("n_obs", C.c_int32), ("status", C.c_int32),
        ("lsmr_itn", C.c_int32), ("lsmr_istop", C.c_int32),
    ]


class HGDesc(C.Structure):
    _fields_ = [("num_stacks", C.c_int), ("num_classes", C.c_int), ("in_h", C.c_int), ("in_w", C.c_int),
                ("max_batch", C.c_int)]


_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t

# name -> (restype, argtypes); mirrors include/df3d_b200.h one to one
SIGNATURES = {
    "df3d_abi_version": (_i, []),
    "df3d_last_error": (C.c_char_p, []),
    "df3d_heatmap_argmax": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "df3d_heatmap_argmax_nhwc": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "df3d_resize_gray_u8": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp]),
    "df3d_jpeg_create": (_i, [C.POINTER(_vp)]),
    "df3d_jpeg_create_backend": (_i, [C.POINTER(_vp), _i]),
    "df3d_jpeg_backend": (_i, [_vp]),
    "df3d_jpeg_destroy": (None, [_vp]),
    "df3d_jpeg_info": (_i, [_vp, _vp, _sz, C.POINTER(_i), C.POINTER(_i)]),
    "df3d_jpeg_decode_gray": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _vp]),
    "df3d_pack_points2d": (_i, [_vp, _i, _i, _i, _i, _i, C.POINTER(C.c_int), _i, _i, _vp, _vp, _vp]),
    "df3d_triangulate_dlt": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "df3d_projection_matrices": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "df3d_bundle_adjust_workspace_bytes": (_sz, [_i, _i, _i]),
    "df3d_bundle_adjust": (_i, [_vp, _vp, _vp, _i, _i, _i, C.POINTER(BAOpts), _vp, _vp, _vp, _sz, _vp]),
    "df3d_bundle_adjust_launches": (_i, [C.POINTER(BAOpts)]),
    "df3d_ba_sharded_plan": (_i, [_i, _i, _i, _i, C.POINTER(_i), C.POINTER(_sz), C.POINTER(_i)]),
    "df3d_ba_sharded_begin": (_i, [_vp, _i, _i, _i, C.POINTER(BAOpts), _vp, _sz, _vp]),
    "df3d_ba_sharded_pass": (_i, [_i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "df3d_ba_sharded_finish": (_i, [_i, _i, _i, _i, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "df3d_ba_sharded_end": (_i, [_vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "df3d_reprojection_error": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "df3d_procrustes_workspace_bytes": (_sz, [_i]),
    "df3d_procrustes": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "df3d_one_euro_filter": (_i, [_vp, _i, _i, C.c_double, C.c_double, C.c_double, C.c_double, _i, _vp, _vp]),
    "df3d_smooth_pose2d": (_i, [_vp, _i, _i, _i, C.c_double, _vp, _vp]),
    "df3d_hg_param_count": (_sz, [C.POINTER(HGDesc)]),
    "df3d_hg_workspace_bytes": (_sz, [C.POINTER(HGDesc)]),
    "df3d_hg_create": (_i, [C.POINTER(HGDesc), _vp, _sz, C.POINTER(_vp)]),
    "df3d_hg_destroy": (None, [_vp]),
    "df3d_hg_forward_argmax": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "df3d_hg_launches_per_forward": (_i, [_vp, _i]),
    "df3d_hg_set_timing": (_i, [_vp, _i]),
    "df3d_hg_read_timing": (_i, [_vp, _vp]),
    "df3d_hg_num_ops": (_i, [_vp]),
    "df3d_hg_op_timing": (_i, [_vp, _i, _vp, _vp, _vp, C.c_char_p, _i]),
    "df3d_hg_set_mean": (_i, [_vp, C.c_float, C.c_float, C.c_float]),
    "df3d_conv2d_nhwc_bf16": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
}


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m deepfly3d_b200.build` "
            "(or __graft_entry__.build()).  deepfly3d_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


class Df3dError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = lib.df3d_last_error()
        raise Df3dError(f"libdf3d_b200 error {rc}: {msg.decode() if msg else ''}")
