"""ctypes binding of libubd.so (include/ubd.h).  No fallback: a missing library is an ImportError,
a missing GPU is a RuntimeError at handle creation."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libubd.so")

UBD_U8, UBD_F32 = 0, 1
PREPROC_NONE, PREPROC_MOBILENET = 0, 1
FP32, TF32, BF16, F16 = 0, 1, 2, 3
PRECISIONS = {"fp32": FP32, "tf32": TF32, "bf16": BF16, "f16": F16}
N_WEIGHT_ARRAYS = 23

STATUS = {0: "UBD_OK", -1: "UBD_ERR_ARG", -2: "UBD_ERR_CUDA", -3: "UBD_ERR_NO_WEIGHTS",
          -4: "UBD_ERR_OVERFLOW", -5: "UBD_ERR_UNSUPPORTED", -6: "UBD_ERR_STATE"}


class Component(C.Structure):
    _fields_ = [("image", C.c_int32), ("label", C.c_int32),
                ("xmin", C.c_int32), ("ymin", C.c_int32), ("xmax", C.c_int32), ("ymax", C.c_int32),
                ("n_pixels", C.c_int32), ("n_filled", C.c_int32), ("area_x2", C.c_int32),
                ("class_id", C.c_int32), ("box", C.c_float * 8)]


COMPONENT_DTYPE = np.dtype([("image", "<i4"), ("label", "<i4"), ("xmin", "<i4"), ("ymin", "<i4"),
                            ("xmax", "<i4"), ("ymax", "<i4"), ("n_pixels", "<i4"), ("n_filled", "<i4"),
                            ("area_x2", "<i4"), ("class_id", "<i4"), ("box", "<f4", (8,))])
assert COMPONENT_DTYPE.itemsize == C.sizeof(Component)

_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
_pp = C.POINTER(C.c_void_p)
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); exactly the symbols include/ubd.h declares
SIGNATURES = {
    "ubd_create": (_i, [_i, _i, _i, _i, _i, _pp]),
    "ubd_destroy": (_i, [_vp]),
    "ubd_last_error": (C.c_char_p, [_vp]),
    "ubd_version": (_i, []),
    "ubd_device_count": (_i, []),
    "ubd_set_weights": (_i, [_vp, _pp, _pi64, _i]),
    "ubd_get_weights": (_i, [_vp, _pp, _pi64, _i]),
    "ubd_set_option": (_i, [_vp, C.c_char_p, _i64]),
    "ubd_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ubd_segment": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "ubd_postprocess": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "ubd_segment_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _i, _vp]),
    "ubd_forward_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ubd_segment_submit": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _i, C.POINTER(C.c_int)]),
    "ubd_segment_submit_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, C.POINTER(C.c_int)]),
    "ubd_segment_wait": (_i, [_vp, _i, _vp, _i, _vp]),
    "ubd_prepare_images": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ubd_prepare_images_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ubd_min_area_box": (_i, [_vp, _i, _vp]),
    "ubd_train_step": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "ubd_train_step_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "ubd_loss": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "ubd_get_grads": (_i, [_vp, _pp, _pi64, _i]),
    "ubd_grad_buffer": (_i, [_vp, _pp, _pi64]),
    "ubd_adam_step": (_i, [_vp, _f, _f, _f, _f, _f]),
    "ubd_train_update": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _f, _f, _f, _f, _vp]),
    "ubd_metric_counts": (_i, [_vp, _vp]),
    "ubd_comm_unique_id": (_i, [_vp]),
    "ubd_comm_init": (_i, [_vp, _vp, _i, _i]),
    "ubd_allreduce_grads": (_i, [_vp]),
    "ubd_comm_destroy": (_i, [_vp]),
    "ubd_debug_dilated_layer": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i]),
    "ubd_debug_wgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ubd_debug_read_trace": (_i, [_vp, _vp, _i]),
    "ubd_synchronize": (_i, [_vp]),
    "ubd_set_stream": (_i, [_vp, _vp]),
    "ubd_get_stat": (_i, [_vp, C.c_char_p, C.POINTER(C.c_double)]),
    "ubd_launch_count": (_i64, [_vp]),
}

_lib = None


def load():
    """Load libubd.so (once).  Raises ImportError with the build hint when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m ubdvss_b200.build` "
            "(ubdvss_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI and the header diverge
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class UbdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


def check(handle, rc):
    if rc != 0:
        msg = load().ubd_last_error(handle)
        raise UbdError(rc, msg.decode() if msg else "")


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
