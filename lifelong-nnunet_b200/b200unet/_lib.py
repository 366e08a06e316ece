"""ctypes binding of libb2unet.so (the C ABI declared in include/b2unet.h).

There is NO CPU fallback: importing this module without the built library raises, and every entry point raises on
a non-zero status with the library's last error string.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libb2unet.so")

B2_F32, B2_BF16 = 0, 1


class B2Error(RuntimeError):
    pass


class Geometry(C.Structure):
    _fields_ = [("batch", C.c_int32), ("in_channels", C.c_int32), ("num_classes", C.c_int32),
                ("base_features", C.c_int32), ("max_features", C.c_int32), ("num_pool", C.c_int32),
                ("patch", C.c_int32 * 3), ("pool", (C.c_int32 * 3) * 7), ("act_dtype", C.c_int32),
                ("lrelu_slope", C.c_float), ("norm_eps", C.c_float)]


class ParamInfo(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("ndim", C.c_int32), ("shape", C.c_int64 * 5), ("numel", C.c_int64)]


class ActView(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("d", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("c", C.c_int32), ("pitch", C.c_int32), ("dtype", C.c_int32)]


class PenEntry(C.Structure):
    _fields_ = [("theta", C.c_void_p), ("theta_star", C.c_void_p), ("fisher", C.c_void_p),
                ("importance", C.c_void_p), ("grad", C.c_void_p), ("numel", C.c_int64)]


class RwEntry(C.Structure):
    _fields_ = [("theta", C.c_void_p), ("grad", C.c_void_p), ("prev", C.c_void_p), ("fisher", C.c_void_p),
                ("score", C.c_void_p), ("numel", C.c_int64)]


class SgdEntry(C.Structure):
    _fields_ = [("theta", C.c_void_p), ("grad", C.c_void_p), ("momentum", C.c_void_p), ("numel", C.c_int64)]


class VitDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("in_channels", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("patch", C.c_int32), ("embed", C.c_int32), ("heads", C.c_int32), ("depth", C.c_int32), ("mlp_ratio", C.c_int32),
                ("out_features", C.c_int32), ("out_c", C.c_int32), ("out_d", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
                ("ln_eps", C.c_float), ("lsa", C.c_int32), ("lsa_mask", C.c_int32)]


class GradBucket(C.Structure):
    _fields_ = [("first_param", C.c_int32), ("event_main", C.c_void_p), ("event_side", C.c_void_p)]


MT_PEN, MT_SGD, MT_RW = 0, 1, 2


class ConvDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("d", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("cin", C.c_int32),
                ("cout", C.c_int32), ("stride", C.c_int32 * 3), ("in_pitch", C.c_int32), ("out_pitch", C.c_int32),
                ("dtype", C.c_int32)]


class AugCase(C.Structure):
    _fields_ = [("volume", C.c_void_p), ("dhw", C.c_int32 * 3), ("lb", C.c_int32 * 3), ("win_lo", C.c_int32 * 3),
                ("win_hi", C.c_int32 * 3)]


class AugSpatial(C.Structure):
    _fields_ = [("m", C.c_float * 9), ("ctr", C.c_float * 3), ("lb", C.c_int32 * 3), ("modified", C.c_int32)]


class AugOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("p", C.c_float * 3)]


class AugBlur(C.Structure):
    _fields_ = [("radius", C.c_int32), ("w", C.c_float * 5)]


AUG_NONE, AUG_NOISE, AUG_MUL, AUG_CONTRAST, AUG_GAMMA_A, AUG_GAMMA_B = range(6)
AUG_MAX_SAMPLES, AUG_MAX_BC, AUG_MAX_SCALES, AUG_MAX_RADIUS = 32, 64, 6, 4


class TconvDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("d", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
                ("k", C.c_int32 * 3), ("in_pitch", C.c_int32), ("out_pitch", C.c_int32), ("dtype", C.c_int32)]


# name -> (restype, argtypes); mirrors include/b2unet.h one to one (tests/test_abi.py checks the export list)
_VP, _I, _F, _I64, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t
SIGNATURES = {
    "b2_version": (_I, []),
    "b2_last_error": (C.c_char_p, []),
    "b2_launch_count": (C.c_longlong, []),
    "b2_set_option": (_I, [C.c_char_p, _I]),
    "b2_unet_plan_create": (_I, [C.POINTER(Geometry), C.POINTER(_VP)]),
    "b2_unet_plan_destroy": (None, [_VP]),
    "b2_unet_num_params": (_I, [_VP]),
    "b2_unet_param_info": (_I, [_VP, _I, C.POINTER(ParamInfo)]),
    "b2_unet_workspace_bytes": (_SZ, [_VP]),
    "b2_unet_output_shape": (_I, [_VP, _I, C.POINTER(C.c_int32 * 3)]),
    "b2_unet_forward": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "b2_unet_backward": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "b2_unet_backward_buckets": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, C.POINTER(GradBucket), _I, _VP]),
    "b2_unet_forward_parts": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "b2_unet_backward_parts": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "b2_unet_head_forward": (_I, [_VP, _VP, _I, _VP, _VP, _VP]),
    "b2_unet_num_convs": (_I, [_VP]),
    "b2_unet_conv_name": (_I, [_VP, _I, C.c_char_p]),
    "b2_unet_conv_output": (_I, [_VP, _VP, _I, C.POINTER(ActView)]),
    "b2_unet_debug_view": (_I, [_VP, _VP, _I, _I, C.POINTER(ActView)]),
    "b2_vit_plan_create": (_I, [C.POINTER(VitDesc), C.POINTER(_VP)]),
    "b2_vit_plan_destroy": (None, [_VP]),
    "b2_vit_workspace_bytes": (_SZ, [_VP]),
    "b2_vit_num_params": (_I, [_VP]),
    "b2_vit_tokens_offset": (_SZ, [_VP]),
    "b2_vit_forward": (_I, [_VP, _VP, C.POINTER(ActView), _VP, C.POINTER(ActView), _VP, _I, _VP]),
    "b2_vit_backward": (_I, [_VP, _VP, C.POINTER(ActView), _VP, _VP, C.POINTER(ActView), _VP, _VP]),
    "b2_dsloss_scratch_bytes": (_SZ, [_I, _I, _I64]),
    "b2_dsloss_fwd_bwd": (_I, [_VP, _VP, _I, _I, _I64, _F, _I, _F, _I, _I, _I, _VP, _VP, _VP, _VP]),
    "b2_quadpen_scratch_bytes": (_SZ, [_I, _I64]),
    "b2_quadpen_fwd_bwd": (_I, [C.POINTER(PenEntry), _I, _F, _VP, _VP, _VP]),
    "b2_multitensor_scratch_bytes": (_SZ, [_I]),
    "b2_fisher_square": (_I, [_VP, _VP, _VP, _I, _VP, _VP]),
    "b2_rw_update": (_I, [C.POINTER(RwEntry), _I, _F, _F, _I, _VP, _VP]),
    "b2_sgd_scratch_bytes": (_SZ, [_I, _I64]),
    "b2_sgd_clip_step": (_I, [C.POINTER(SgdEntry), _I, _F, _F, _F, _I, _F, _I, _VP, _VP, _VP]),
    "b2_mt_blob_bytes": (_SZ, [_I, _I]),
    "b2_mt_blob_build": (_I, [_I, _VP, _I, _VP, C.POINTER(C.c_int32)]),
    "b2_mt_part_bytes": (_SZ, [_I]),
    "b2_quadpen_dev": (_I, [_VP, _I, _I, _F, _VP, _VP, _VP]),
    "b2_sgd_clip_step_dev": (_I, [_VP, _I, _I, _VP, _I, _VP, _VP, _VP]),
    "b2_rw_update_dev": (_I, [_VP, _I, _I, _F, _F, _I, _VP]),
    "b2_kd_scratch_bytes": (_SZ, [_I, _I, _I64]),
    "b2_kd_lwf": (_I, [_VP, _VP, _I, _I, _I64, _F, _VP, _VP, _VP]),
    "b2_kd_mib": (_I, [_VP, _VP, _I, _I, _I64, _F, _F, _VP, _VP, _VP, _VP]),
    "b2_plop_pseudo": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _VP, _F, _F, _VP, _VP, _VP, _VP]),
    "b2_plop_entropy_hist": (_I, [_VP, _VP, _I, _I, _I64, _F, _I, _VP, _VP]),
    "b2_pod_scratch_bytes": (_SZ, [C.POINTER(ActView), _I]),
    "b2_pod_local": (_I, [C.POINTER(ActView), C.POINTER(ActView), _I, _VP, _VP, _VP]),
    "b2_online_eval": (_I, [_VP, _VP, _I, _I, _I64, _VP, _VP, _VP]),
    "b2_sliding_accumulate": (_I, [_VP, _I, _I, _I, _I, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _F, _I, _VP]),
    "b2_sliding_finalize": (_I, [_VP, _VP, _I, _I64, _VP, _VP]),
    "b2_aug_crop": (_I, [C.POINTER(AugCase), _I, _I, C.POINTER(C.c_int32 * 3), _VP, _VP, _VP]),
    "b2_aug_spatial": (_I, [C.POINTER(AugSpatial), _I, _I, C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3), _VP, _VP, _VP, _VP, _VP]),
    "b2_aug_stats_scratch_bytes": (_SZ, [_I, _I64]),
    "b2_aug_stats": (_I, [_VP, _I, _I64, _VP, _VP, _VP]),
    "b2_aug_pointwise": (_I, [C.POINTER(AugOp), _I, _I64, _VP, _VP, _VP, C.c_uint64, _VP]),
    "b2_aug_blur": (_I, [C.POINTER(AugBlur), _I, C.POINTER(C.c_int32 * 3), _VP, _VP, _VP]),
    "b2_aug_lowres_scratch_bytes": (_SZ, [C.POINTER(C.c_int32 * 3)]),
    "b2_aug_lowres": (_I, [_VP, C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3), _VP, _VP]),
    "b2_aug_finalize": (_I, [C.POINTER(C.c_int32), _I, _I, C.POINTER(C.c_int32 * 3), _VP, _VP, _VP, C.POINTER(_VP), C.POINTER(C.c_int32), _I, _VP]),
    "b2_conv3d_scratch_bytes": (_SZ, [C.POINTER(ConvDesc)]),
    "b2_conv3d_fwd": (_I, [C.POINTER(ConvDesc), _VP, _VP, _VP, _VP, _VP, _F, _VP, _VP]),
    "b2_conv3d_shadow_bytes": (_SZ, [C.POINTER(ConvDesc)]),
    "b2_conv3d_make_shadow": (_I, [C.POINTER(ConvDesc), _VP, _VP, _VP]),
    "b2_conv3d_fwd_shadow": (_I, [C.POINTER(ConvDesc), _VP, _VP, _VP, _VP, _VP, _VP]),
    "b2_conv3d_fwd_shadow_stats": (_I, [C.POINTER(ConvDesc), _VP, _VP, _VP, _VP, _VP, _VP]),
    "b2_conv3d_bwd": (_I, [C.POINTER(ConvDesc), _VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP]),
    "b2_tconv3d_scratch_bytes": (_SZ, [C.POINTER(TconvDesc)]),
    "b2_tconv3d_fwd": (_I, [C.POINTER(TconvDesc), _VP, _VP, _VP, _VP, _VP]),
    "b2_tconv3d_bwd": (_I, [C.POINTER(TconvDesc), _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "b2_norm_scratch_bytes": (_SZ, [_I, _I64, _I]),
    "b2_norm_lrelu_fwd": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I64, _I, _I, _I, _I, _F, _VP]),
    "b2_norm_lrelu_bwd": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I64, _I, _I, _I, _I, _I, _I, _F, _VP, _VP]),
}

_lib = None


def load():
    """Load libb2unet.so (once).  Raises B2Error when the library has not been built -- by design there is no
    fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2Error("libb2unet.so not found at %s -- build it with `python -c \"import __graft_entry__ as g; "
                      "g.build()\"` (make -C lifelong-nnunet_b200/csrc)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing: loud by design
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    # tuning switches for experiments: B2_OPTIONS="name=value,name=value" (see b2_set_option in csrc/plan.cu)
    for kv in filter(None, os.environ.get("B2_OPTIONS", "").split(",")):
        name, _, val = kv.partition("=")
        rc = lib.b2_set_option(name.strip().encode(), int(val))
        if rc != 0:
            raise B2Error("B2_OPTIONS: %s" % lib.b2_last_error().decode())
    return lib


def check(rc):
    if rc != 0:
        raise B2Error("libb2unet error %d: %s" % (rc, load().b2_last_error().decode()))


def launch_count():
    return int(load().b2_launch_count())
