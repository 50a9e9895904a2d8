"""ctypes declarations for libacf_b200.so (include/acf_b200.h).

The library is the product; this module only declares its C ABI for Python callers
(tests, bench.py).  Importing it fails loudly if the shared library has not been built --
there is no Python or CPU fallback path.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACFB_LIB") or os.path.join(HERE, "libacf_b200.so")  # ACFB_LIB: A/B another build of the library


class Options(C.Structure):
    _fields_ = [
        ("shrink", C.c_int32), ("color_enabled", C.c_int32), ("color_smooth", C.c_double), ("color_space", C.c_int32),
        ("gm_enabled", C.c_int32), ("gm_colorChn", C.c_int32), ("gm_normRad", C.c_int32), ("gm_normConst", C.c_double),
        ("gm_full", C.c_int32),
        ("gh_enabled", C.c_int32), ("gh_binSize", C.c_int32), ("gh_nOrients", C.c_int32), ("gh_softBin", C.c_int32),
        ("gh_useHog", C.c_int32), ("gh_clipHog", C.c_double),
        ("nPerOct", C.c_int32), ("nOctUp", C.c_int32), ("nApprox", C.c_int32),
        ("nLambdas", C.c_int32), ("lambdas", C.c_double * 8),
        ("pad_w", C.c_int32), ("pad_h", C.c_int32), ("minDs_w", C.c_int32), ("minDs_h", C.c_int32),
        ("smooth", C.c_double), ("concat", C.c_int32),
        ("modelDs_w", C.c_int32), ("modelDs_h", C.c_int32), ("modelDsPad_w", C.c_int32), ("modelDsPad_h", C.c_int32),
        ("stride", C.c_int32), ("cascThr", C.c_double), ("cascCal", C.c_double),
        ("nms_type", C.c_char * 16), ("nms_overlap", C.c_double), ("nms_ovrDnm", C.c_char * 16),
    ]


class Classifier(C.Structure):
    _fields_ = [("nTrees", C.c_int32), ("nTreeNodes", C.c_int32), ("treeDepth", C.c_int32),
                ("fids", C.c_void_p), ("thrs", C.c_void_p), ("child", C.c_void_p), ("hs", C.c_void_p),
                ("weights", C.c_void_p), ("depth", C.c_void_p)]


class Det(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("score", C.c_float),
                ("frame", C.c_int32)]


class Hit(C.Structure):
    _fields_ = [("frame", C.c_int32), ("scale", C.c_int32), ("c", C.c_int32), ("r", C.c_int32), ("score", C.c_float)]


class Channels(C.Structure):
    _fields_ = [("data", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("nchn", C.c_int32),
                ("scale", C.c_double), ("scalehw_w", C.c_double), ("scalehw_h", C.c_double)]


class ScaleInfo(C.Structure):
    _fields_ = [("scale", C.c_double), ("scalehw_w", C.c_double), ("scalehw_h", C.c_double),
                ("h", C.c_int32), ("w", C.c_int32), ("pitch", C.c_int32), ("nchn", C.c_int32), ("is_real", C.c_int32),
                ("real_index", C.c_int32), ("offset", C.c_int64)]


class Modify(C.Structure):  # acfb_modify
    _fields_ = [("has_nPerOct", C.c_int32), ("nPerOct", C.c_int32), ("has_nOctUp", C.c_int32), ("nOctUp", C.c_int32),
                ("has_nApprox", C.c_int32), ("nApprox", C.c_int32), ("has_lambdas", C.c_int32), ("nLambdas", C.c_int32),
                ("lambdas", C.c_double * 8), ("has_pad", C.c_int32), ("pad_w", C.c_int32), ("pad_h", C.c_int32),
                ("has_minDs", C.c_int32), ("minDs_w", C.c_int32), ("minDs_h", C.c_int32), ("has_nms", C.c_int32),
                ("nms_type", C.c_char * 16), ("nms_overlap", C.c_double), ("nms_ovrDnm", C.c_char * 16),
                ("has_stride", C.c_int32), ("stride", C.c_int32), ("has_cascThr", C.c_int32), ("cascThr", C.c_double), ("cascCal", C.c_double)]


# every symbol include/acf_b200.h declares: (restype, argtypes)
_vp, _i, _sz, _d = C.c_void_p, C.c_int, C.c_size_t, C.c_double
_pi = C.POINTER(C.c_int)
SYMBOLS = {
    "acfb_last_error": (C.c_char_p, []),
    "acfb_version": (C.c_char_p, []),
    "acfb_model_load": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "acfb_model_load_file": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "acfb_model_create": (_i, [C.POINTER(Options), C.POINTER(Classifier), C.POINTER(_vp)]),
    "acfb_model_save": (_i, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "acfb_model_save_file": (_i, [_vp, C.c_char_p]),
    "acfb_model_options": (_i, [_vp, C.POINTER(Options)]),
    "acfb_model_classifier": (_i, [_vp, C.POINTER(Classifier)]),
    "acfb_model_modify": (_i, [_vp, _d, _d, _i]),
    "acfb_model_modify_ex": (_i, [_vp, _vp]),
    "acfb_model_destroy": (None, [_vp]),
    "acfb_engine_create": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_vp)]),
    "acfb_engine_destroy": (None, [_vp]),
    "acfb_set_nms": (_i, [_vp, _i]),
    "acfb_set_max_detection_count": (_i, [_vp, _i]),
    "acfb_set_detection_score_prune_ratio": (_i, [_vp, _d]),
    "acfb_set_input_format": (_i, [_vp, _i]),
    "acfb_set_is_transpose": (_i, [_vp, _i]),
    "acfb_set_is_luv": (_i, [_vp, _i]),
    "acfb_set_hit_capacity": (_i, [_vp, _i]),
    "acfb_get_scales": (_i, [C.POINTER(Options), _i, _i, C.POINTER(_d), C.POINTER(_d), _i, _pi]),
    "acfb_plan": (_i, [_vp, _i, _i, C.POINTER(ScaleInfo), _i, _pi, C.POINTER(C.c_int64)]),
    "acfb_pyramid": (_i, [_vp, _vp, _i, _i, _i, _i]),
    "acfb_pyramid_device_ptr": (_i, [_vp, _i, C.POINTER(_vp)]),
    "acfb_pyramid_read": (_i, [_vp, _i, _i, _vp, _sz]),
    "acfb_pyramid_lambdas": (_i, [_vp, C.POINTER(_d), _i, _pi]),
    "acfb_detect_pyramid": (_i, [_vp, C.POINTER(Det), _i, _pi, _pi]),
    "acfb_detect": (_i, [_vp, _vp, _i, _i, _i, _i, C.POINTER(Det), _i, _pi, _pi]),
    "acfb_last_hits": (_i, [_vp, C.POINTER(Hit), _i, _pi, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "acfb_acf_detect1": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _pi, C.POINTER(C.c_uint64)]),
    "acfb_acf_detect1_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _pi, C.POINTER(C.c_uint64)]),
    "acfb_detect_channels": (_i, [_vp, C.POINTER(Channels), _i, _i, C.POINTER(Det), _i, _pi]),
    "acfb_evaluate": (_i, [_vp, _vp, _i, _i, C.POINTER(C.c_float)]),
    "acfb_submit": (_i, [_vp, _vp, _i, _i, _i, _i]),
    "acfb_collect": (_i, [_vp, C.POINTER(Det), _i, _pi, _pi]),
    "acfb_synchronize": (_i, [_vp]),
    "acfb_dist_unique_id": (_i, [_vp]),
    "acfb_dist_init_rank": (_i, [_vp, _vp, _i, _i]),
    "acfb_dist_init_all": (_i, [C.POINTER(_vp), _i]),
    "acfb_dist_collect": (_i, [_vp, C.POINTER(Det), _i, _pi, _pi]),
    "acfb_dist_info": (_i, [_vp, _pi, _pi, _pi]),
    "acfb_selftest_exchange": (_i, [_i, _i, _i]),
    "acfb_selftest_math": (_i, [_vp, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]),
    "acfb_launch_count": (C.c_uint64, [_vp]),
    "acfb_stream": (C.c_uint64, [_vp]),
    "acfb_stage_times": (_i, [_vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), _i]),
    "acfb_enable_stage_timing": (_i, [_vp, _i]),
    "acfb_collect_times": (_i, [_vp, C.POINTER(_d), C.POINTER(_d)]),
    "acfb_compute_channels": (_i, [_vp, _vp, _i, _i, _vp, _sz, _pi, _pi, _pi]),
    "acfb_op_rgb_convert": (_i, [_vp, _vp, _i, _i, _i, _vp, _pi]),
    "acfb_op_conv_tri": (_i, [_vp, _vp, _i, _i, _i, C.c_double, _vp]),
    "acfb_op_gradient_mag": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.c_double, _i, _vp, _vp]),
    "acfb_op_gradient_hist": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, C.c_double, _i, _vp]),
    "acfb_op_im_resample": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.c_double, _vp]),
    "acfb_set_debug_taps": (_i, [_vp, _i]),
    "acfb_tap": (_i, [_vp, C.c_char_p, _i, _i, _vp, _sz, _pi, _pi, _pi]),
}

_lib = None


def lib():
    """Load libacf_b200.so (once).  Raises if it is missing: build it with `make -C acf_b200`."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built -- run `make -C acf_b200` (or __graft_entry__.build()); "
                              "acf_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class AcfError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise AcfError(lib().acfb_last_error().decode(errors="replace"))
