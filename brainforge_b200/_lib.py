"""ctypes binding of libbf_b200.so (the C ABI declared in include/bf_b200.h).

The product path has NO fallback: if the shared library is missing or a symbol is absent this module
raises at import time, and every wrapper raises when the library reports an error.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbf_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "bf_b200.h")

# enums (mirrors of include/bf_b200.h)
OK, ERR_CUDA, ERR_VALUE, ERR_ARG, ERR_WORKSPACE = 0, 1, 2, 3, 4
ACT = {"linear": 0, "sigmoid": 1, "tanh": 2, "relu": 3, "softmax": 4}
COST = {"mse": 0, "mean_squared_error": 0, "cxent": 1, "categorical_crossentropy": 1, "bxent": 2, "binary_crossentropy": 2,
        "hinge": 3}
OPT = {"sgd": 0, "momentum": 1, "nesterov": 2, "adagrad": 3, "rmsprop": 4, "adam": 5}
CONV_MODE = {"valid": 0, "full": 1}
PREC = {"fp32": 0, "tf32": 1}

_c = ctypes
_vp, _i, _f, _sz = _c.c_void_p, _c.c_int, _c.c_float, _c.c_size_t

# name -> (restype, argtypes).  Kept explicit so that a signature drift against the header is a test failure.
SIGNATURES = {
    "bf_last_error": (_c.c_char_p, []),
    "bf_version": (_i, []),
    "bf_device_info": (_i, [_c.POINTER(_i)] * 3),
    "bf_set_workspace": (_i, [_vp, _sz]),
    "bf_last_path": (_i, []),
    "bf_launch_count": (_c.c_ulonglong, []),
    "bf_memcpy_h2d": (_i, [_vp, _vp, _sz, _vp]),
    "bf_upload_host": (_i, [_vp, _vp, _sz, _i, _vp]),
    "bf_memcpy_d2h": (_i, [_vp, _vp, _sz, _vp]),
    "bf_memcpy_d2d": (_i, [_vp, _vp, _sz, _vp]),
    "bf_stream_sync": (_i, [_vp]),
    "bf_cast_f64_to_f32": (_i, [_vp, _vp, _sz, _vp]),
    "bf_cast_f32_to_f64": (_i, [_vp, _vp, _sz, _vp]),
    "bf_transpose01": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "bf_affine": (_i, [_vp, _vp, _sz, _f, _f, _vp]),
    "bf_mul_rowmask": (_i, [_vp, _vp, _vp, _sz, _sz, _f, _vp]),
    "bf_act_forward": (_i, [_i, _vp, _vp, _sz, _vp]),
    "bf_softmax_forward": (_i, [_vp, _vp, _i, _i, _vp]),
    "bf_act_backward": (_i, [_i, _vp, _vp, _vp, _sz, _vp]),
    "bf_dense_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "bf_dense_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "bf_dense_tma_supported": (_i, [_i] * 4),
    "bf_dense_head_supported": (_i, [_i] * 4),
    "bf_mlp2_step_supported": (_i, [_i] * 7),
    "bf_mlp2_step": (_i, [_vp] * 15 + [_i] * 7 + [_vp]),
    "bf_dense_head": (_i, [_vp] * 11 + [_i] * 8 + [_vp]),
    "bf_dense_forward_cl": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "bf_dense_backward_cl": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "bf_conv_forward": (_i, [_vp, _vp, _vp, _vp] + [_i] * 11 + [_vp]),
    "bf_conv_backward": (_i, [_vp] * 6 + [_i] * 8 + [_vp]),
    "bf_conv_nhwc_supported": (_i, [_i] * 4),
    "bf_conv_nhwc_forward": (_i, [_vp] * 4 + [_i] * 8 + [_vp]),
    "bf_conv_nhwc_backward_data": (_i, [_vp] * 3 + [_i] * 7 + [_vp]),
    "bf_conv_nhwc_backward_filter": (_i, [_vp] * 4 + [_i] * 7 + [_vp]),
    "bf_conv_c3_supported": (_i, [_i] * 6),
    "bf_conv_c3_forward": (_i, [_vp] * 4 + [_i] * 8 + [_vp]),
    "bf_conv_c3_forward_pool": (_i, [_vp] * 5 + [_i] * 9 + [_vp]),
    "bf_conv_c3_backward_filter": (_i, [_vp] * 4 + [_i] * 7 + [_vp]),
    "bf_conv_c3_backward_filter_unpool": (_i, [_vp] * 6 + [_i] * 7 + [_vp]),
    "bf_pool_mask_bytes": (_sz, [_i] * 5),
    "bf_pool_forward": (_i, [_vp, _vp, _vp] + [_i] * 5 + [_vp]),
    "bf_pool_backward": (_i, [_vp, _vp, _vp] + [_i] * 5 + [_vp]),
    "bf_pool_mask_expand": (_i, [_vp, _vp] + [_i] * 5 + [_vp]),
    "bf_layout_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "bf_layout_nhwc_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "bf_pool_nhwc_forward": (_i, [_vp, _vp, _vp] + [_i] * 6 + [_vp]),
    "bf_pool_nhwc_backward": (_i, [_vp, _vp, _vp, _vp] + [_i] * 5 + [_vp]),
    "bf_pool_nhwc_db_rows": (_i, [_i] * 5),
    "bf_pool_nhwc_backward_db": (_i, [_vp] * 5 + [_i] * 5 + [_vp]),
    "bf_conv_nhwc_repack": (_i, [_vp] * 3 + [_i] * 4 + [_vp]),
    "bf_conv_nhwc_forward_packed": (_i, [_vp] * 4 + [_i] * 8 + [_vp]),
    "bf_conv_nhwc_forward_pool_packed": (_i, [_vp] * 5 + [_i] * 8 + [_vp]),
    "bf_conv_nhwc_backward_data_packed": (_i, [_vp] * 3 + [_i] * 7 + [_vp]),
    "bf_conv_nhwc_backward_filter_dbp": (_i, [_vp] * 5 + [_i] * 8 + [_vp]),
    "bf_gap_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "bf_gap_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "bf_cost_delta": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "bf_cost_value": (_i, [_i, _vp, _vp, _i, _i, _vp, _vp]),
    "bf_cost_delta_value": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "bf_accuracy": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "bf_lstm_forward": (_i, [_vp] * 7 + [_i] * 6 + [_vp]),
    "bf_lstm_forward_bm": (_i, [_vp] * 7 + [_i] * 6 + [_vp]),
    "bf_lstm_backward": (_i, [_vp] * 9 + [_i] * 6 + [_vp]),
    "bf_rnn_forward": (_i, [_vp] * 5 + [_i] * 6 + [_vp]),
    "bf_rnn_backward": (_i, [_vp] * 7 + [_i] * 6 + [_vp]),
    "bf_clockwork_forward": (_i, [_vp] * 6 + [_i] * 6 + [_vp]),
    "bf_clockwork_backward": (_i, [_vp] * 8 + [_i] * 6 + [_vp]),
    "bf_gru_forward": (_i, [_vp] * 7 + [_i] * 6 + [_vp]),
    "bf_gru_backward": (_i, [_vp] * 8 + [_i] * 6 + [_vp]),
    "bf_opt_step": (_i, [_i, _vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _f, _i, _vp]),
    "bf_opt_step_allreduce": (_i, [_i, _vp, _vp, _vp, _vp, _sz, _sz, _f, _f, _f, _f, _f, _i, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _vp]),
    "bf_fill": (_i, [_vp, _sz, _f, _vp]),
    "bf_dense_cl_supported": (_i, [_i, _i]),
    "bf_fitness_mlp": (_i, [_vp] * 4 + [_i] * 6 + [_vp]),
    "bf_population_fitness": (_i, [_vp, _c.c_long, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "bf_population_relayout": (_i, [_vp, _c.c_long, _vp, _vp, _i, _i, _i, _vp]),
    "bf_population_next_generation": (_i, [_vp, _vp, _c.c_long, _i, _i, _vp, _vp, _vp, _f, _c.c_uint, _c.c_uint, _vp, _vp, _vp,
                                           _vp, _vp]),
    "bf_gradcheck_perturb": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _i, _vp]),
    "bf_gradcheck_cost": (_i, [_i, _vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "bf_gather_rows": (_i, [_vp, _vp, _vp, _i, _sz, _vp]),
}


def header_symbols(path=HEADER):
    """Every function name include/bf_b200.h declares."""
    text = open(path).read()
    return sorted(set(re.findall(r"BF_API\s+[\w\s\*]+?\b(bf_\w+)\s*\(", text)))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "brainforge_b200: %s is missing. Build it with `python brainforge_b200/build.py` "
            "(there is no CPU or PyTorch fallback for this path)." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError("brainforge_b200: libbf_b200.so does not export %s" % name) from e
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class WorkspaceTooSmall(RuntimeError):
    def __init__(self, msg, need):
        super().__init__(msg)
        self.need = need


class Unsupported(RuntimeError):
    """BF_ERR_ARG: the kernel family asked for cannot take this shape (callers may pick another family)."""


def last_error():
    return (lib.bf_last_error() or b"").decode("utf-8", "replace")


def check(rc):
    """Map the library's status codes onto the reference's exceptions (SURVEY 8b error convention)."""
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_VALUE:
        raise ValueError(msg)
    if rc == ERR_WORKSPACE:
        m = re.search(r"need (\d+) bytes", msg)
        raise WorkspaceTooSmall(msg, int(m.group(1)) if m else 0)
    if rc == ERR_ARG:
        raise Unsupported("libbf_b200: " + msg)
    raise RuntimeError("libbf_b200: " + msg)
