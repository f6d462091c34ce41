"""ctypes binding of libtuber_b200.so -- the C-ABI declared in include/tuber_b200.h.

The library is the product: there is no Python or CPU fallback.  Importing this module without
the built library raises, and creating a plan without an sm_100 device fails inside the library.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtuber_b200.so")

TUBER_ABI_VERSION = 1
NUM_STAGES = 10
POOL = {"avg": 0, "max": 1, "decode": 2, "center": 3, "none": 4}
FMT_F32, FMT_SPLIT = 0, 1
TUBER_OK, TUBER_ERR_INVALID, TUBER_ERR_MISSING, TUBER_ERR_SHAPE, TUBER_ERR_CUDA, TUBER_ERR_STATE = 0, -1, -2, -3, -4, -5


class TuberConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("blocks", C.c_int32 * 4), ("last_stride", C.c_int32),
                ("pool", C.c_int32), ("pool_kernel", C.c_int32), ("d_model", C.c_int32), ("nhead", C.c_int32),
                ("enc_layers", C.c_int32), ("dec_layers", C.c_int32), ("dim_ff", C.c_int32),
                ("num_queries", C.c_int32), ("num_classes", C.c_int32), ("ava_mode", C.c_int32)]


class TuberShapeInfo(C.Structure):
    _fields_ = [("Tf", C.c_int32), ("Hf", C.c_int32), ("Wf", C.c_int32), ("Tp", C.c_int32),
                ("enc_tokens", C.c_int32), ("cls_tokens", C.c_int32), ("launches", C.c_int32),
                ("workspace_bytes", C.c_int64)]


class TuberKernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int32), ("ms", C.c_float), ("bytes", C.c_double),
                ("flops", C.c_double)]


class TuberError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"tuber_b200 error {status}: {message}")
        self.status = status


_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
# name -> (restype, argtypes); every entry point of include/tuber_b200.h
PROTOTYPES = {
    "tuber_abi_version": (_I, []),
    "tuber_last_error": (C.c_char_p, []),
    "tuber_plan_create": (_I, [C.POINTER(TuberConfig), C.POINTER(_P)]),
    "tuber_plan_destroy": (None, [_P]),
    "tuber_plan_set_weight": (_I, [_P, C.c_char_p, _P, C.POINTER(_L), _I]),
    "tuber_plan_finalize": (_I, [_P]),
    "tuber_forward": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "tuber_forward_host": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "tuber_forward_host_submit": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "tuber_forward_host_wait": (_I, [_P, _I]),
    "tuber_input_lut": (_I, [C.POINTER(_F), C.POINTER(_F), C.POINTER(_F)]),
    "tuber_set_input_norm": (_I, [_P, C.POINTER(_F), C.POINTER(_F)]),
    "tuber_forward_u8": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "tuber_forward_host_u8": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "tuber_forward_host_u8_submit": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "tuber_has_ltc": (_I, [_P]),
    "tuber_forward_ltc": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _I, _I, _P, _P, _P, _P, _P]),
    "tuber_query_shapes": (_I, [_P, _I, _I, _I, _I, C.POINTER(TuberShapeInfo)]),
    "tuber_set_graph": (_I, [_P, _I]),
    "tuber_set_force_simt": (_I, [_P, _I]),
    "tuber_set_debug_keep": (_I, [_P, _I]),
    "tuber_last_launches": (_I, [_P]),
    "tuber_graph_count": (_I, [_P]),
    "tuber_set_profiling": (_I, [_P, _I]),
    "tuber_get_stage_ms": (_I, [_P, C.POINTER(_F)]),
    "tuber_get_stage_work": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "tuber_stage_name": (C.c_char_p, [_I]),
    "tuber_set_kernel_profiling": (_I, [_P, _I]),
    "tuber_get_kernel_profile": (_I, [_P, C.POINTER(TuberKernelStat), _I, C.POINTER(_I)]),
    "tuber_debug_fetch": (_I, [_P, C.c_char_p, _P, C.POINTER(_L), _P]),
    "tuber_postprocess": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P]),
    "tuber_op_to_split": (_I, [_P, _P, _L, _I, _P]),
    "tuber_op_from_split": (_I, [_P, _P, _L, _I, _P]),
    "tuber_op_pack_weight": (_I, [_P, _P, _I, _I, _P]),
    "tuber_op_gemm_tc": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _P]),
    "tuber_op_gemm_tc_fused2": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P]),
    "tuber_op_sgemm": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "tuber_op_dwconv": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "tuber_op_stem": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "tuber_op_layernorm": (_I, [_P, _P, _P, _P, _P, _L, _I, _P]),
    "tuber_op_attention": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P]),
    "tuber_op_attention_kernel": (C.c_char_p, [_I, _I, _I, _I, _I, _I]),
    "tuber_op_normalize_u8": (_I, [_P, C.POINTER(_F), C.POINTER(_F), _P, _I, _L, _P]),
    "tuber_op_posenc": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "tuber_frames_create": (_I, [C.POINTER(_P), _I]),
    "tuber_frames_destroy": (None, [_P]),
    "tuber_frames_decode": (_I, [_P, C.POINTER(_P), C.POINTER(_L), _I, _I, _I, _P, _P]),
    "tuber_frames_last_error": (C.c_char_p, []),
    "tuber_op_jpeg_coefficients": (_I, [_P, _L, _P, _L, C.POINTER(_I)]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library (built by ``build.py`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (or tubelet-transformer_b200/build.py) "
                          "to compile the sm_100a library; tuber_b200 has no fallback path")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)           # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.tuber_abi_version() != TUBER_ABI_VERSION:
        raise ImportError("libtuber_b200.so ABI version mismatch")
    _lib = lib
    return lib


def input_lut(mean, std):
    """The 3 x 256 value table of the reference's ToTensor + Normalize (host-only; tuber_input_lut) as a list of 768 floats."""
    m, s, out = (_F * 3)(*mean), (_F * 3)(*std), (_F * 768)()
    check(load().tuber_input_lut(m, s, out))
    return list(out)


def check(status: int) -> None:
    if status != 0:
        raise TuberError(status, load().tuber_last_error().decode("utf-8", "replace"))
