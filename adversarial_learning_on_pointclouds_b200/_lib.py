"""ctypes binding of libpcadv.so (the C ABI declared in include/pcadv.h).

The library is built in-tree by ``_build.build()``.  There is no CPU or PyTorch
fallback: if the shared object is missing or the device is not sm_100,
``lib()`` raises and every op fails loudly.
"""
import ctypes as C
import os

from . import _build

MAX_SEG = 6
F32, F16, BF16, I32, I64 = 0, 1, 2, 3, 4
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
ENGINE_SIMT, ENGINE_TC = 0, 1


class Seg(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int64), ("k", C.c_int32), ("dtype", C.c_int32)]


class LinearArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("n", C.c_int32), ("num_seg", C.c_int32),
        ("seg", Seg * MAX_SEG),
        ("w", C.c_void_p), ("ldw", C.c_int64), ("w_dtype", C.c_int32), ("engine", C.c_int32),
        ("bias", C.c_void_p), ("group_bias", C.c_void_p), ("rows_per_group", C.c_int64),
        ("addend", C.c_void_p), ("ld_addend", C.c_int64),
        ("act", C.c_int32), ("slope", C.c_float),
        ("mask", C.c_void_p), ("ld_mask", C.c_int64), ("mask_dtype", C.c_int32),
        ("mask_act", C.c_int32), ("mask_slope", C.c_float), ("out_dtype", C.c_int32),
        ("out_scale", C.c_void_p), ("out", C.c_void_p), ("ld_out", C.c_int64),
        ("colmax_key", C.c_void_p), ("rowmax_key", C.c_void_p),
        ("bits_out", C.c_void_p), ("ld_bits_out", C.c_int64),
        ("mask_bits", C.c_void_p), ("ld_mask_bits", C.c_int64),
        ("seg0_group_sum", C.c_void_p),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("n", C.c_int32), ("num_seg", C.c_int32),
        ("dz", C.c_void_p), ("ld_dz", C.c_int64), ("dz_dtype", C.c_int32), ("engine", C.c_int32),
        ("seg", Seg * MAX_SEG),
        ("dw", C.c_void_p), ("ld_dw", C.c_int64), ("dbias", C.c_void_p),
        ("dgroup_bias", C.c_void_p), ("rows_per_group", C.c_int64), ("scale", C.c_void_p),
    ]


class MaxBwdArgs(C.Structure):
    _fields_ = [
        ("groups", C.c_int32), ("n", C.c_int32), ("k", C.c_int32), ("act", C.c_int32),
        ("slope", C.c_float), ("x_dtype", C.c_int32), ("w_dtype", C.c_int32), ("prev_act", C.c_int32),
        ("prev_slope", C.c_float), ("dz_dtype", C.c_int32), ("dz_inout", C.c_void_p),
        ("ld_dz", C.c_int64), ("workspace", C.c_void_p), ("rows_per_group", C.c_int64),
        ("dg", C.c_void_p), ("gval", C.c_void_p), ("idx", C.c_void_p),
        ("x", C.c_void_p), ("ldx", C.c_int64), ("w", C.c_void_p), ("ldw", C.c_int64),
        ("dw", C.c_void_p), ("ld_dw", C.c_int64), ("dbias", C.c_void_p),
        ("dx_acc", C.c_void_p), ("ld_dx", C.c_int64), ("scale", C.c_void_p),
    ]


class HeadArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("n", C.c_int32), ("mode", C.c_int32),
        ("logits", C.c_void_p), ("ld", C.c_int64), ("labels", C.c_void_p),
        ("probs", C.c_void_p), ("ld_probs", C.c_int64), ("probs_dtype", C.c_int32),
        ("probs_cols", C.c_int32),
        ("dz", C.c_void_p), ("ld_dz", C.c_int64), ("dz_dtype", C.c_int32), ("dz_cols", C.c_int32),
        ("dz_gain", C.c_float), ("loss_sum", C.c_void_p), ("valid_count", C.c_void_p),
    ]


class ChainLayer(C.Structure):
    _fields_ = [("w", C.c_void_p), ("ldw", C.c_int64), ("n", C.c_int32), ("act", C.c_int32),
                ("slope", C.c_float), ("bias", C.c_void_p), ("out", C.c_void_p), ("ld_out", C.c_int64),
                ("bits_out", C.c_void_p), ("ld_bits", C.c_int64)]


class ChainArgs(C.Structure):
    _fields_ = [("rows", C.c_int64), ("x", C.c_void_p), ("ldx", C.c_int64), ("k0", C.c_int32),
                ("dtype", C.c_int32), ("num_layers", C.c_int32), ("layer", ChainLayer * 4),
                ("rowmax_key", C.c_void_p), ("out_f32", C.c_void_p), ("n_f32", C.c_int32)]


class BackLevelArgs(C.Structure):
    _fields_ = [("rows", C.c_int64), ("n", C.c_int32), ("num_seg", C.c_int32), ("seg", Seg * MAX_SEG),
                ("w", C.c_void_p), ("ldw", C.c_int64), ("x", C.c_void_p), ("ldx", C.c_int64),
                ("mask_bits", C.c_void_p), ("ld_mask_bits", C.c_int64), ("mask_act", C.c_int32),
                ("mask_slope", C.c_float), ("dz_out", C.c_void_p), ("ld_out", C.c_int64),
                ("dw", C.c_void_p * MAX_SEG), ("ld_dw", C.c_int64 * MAX_SEG),
                ("dbias", C.c_void_p * MAX_SEG), ("dgroup", C.c_void_p * MAX_SEG),
                ("rows_per_group", C.c_int64), ("scale", C.c_void_p),
                ("onehot_dy", C.c_void_p), ("onehot_val", C.c_void_p), ("onehot_idx", C.c_void_p),
                ("onehot_act", C.c_int32), ("onehot_slope", C.c_float), ("onehot_scale", C.c_void_p)]


HEAD_CE, HEAD_LSM = 0, 1
WS_MAXPOOL_BWD_INPLACE, WS_AMAX_SCALE = 0, 1

# every symbol include/pcadv.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pcadv_linear": (C.c_int, [C.POINTER(LinearArgs), C.c_void_p]),
    "pcadv_wgrad": (C.c_int, [C.POINTER(WgradArgs), C.c_void_p]),
    "pcadv_chain": (C.c_int, [C.POINTER(ChainArgs), C.c_void_p]),
    "pcadv_backlevel": (C.c_int, [C.POINTER(BackLevelArgs), C.c_void_p]),
    "pcadv_max_finalize": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "pcadv_maxpool_bwd": (C.c_int, [C.POINTER(MaxBwdArgs), C.c_void_p]),
    "pcadv_rowmax_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                   C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_int32, C.c_void_p]),
    "pcadv_rowmax_dgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                     C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                     C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_float,
                                     C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "pcadv_rowmax_wgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_float, C.c_void_p, C.c_int64, C.c_int32,
                                     C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pcadv_amax_scale": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_float,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "pcadv_convert": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
                                C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
    "pcadv_convert_cm": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                   C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "pcadv_softmax_head": (C.c_int, [C.POINTER(HeadArgs), C.c_void_p]),
    "pcadv_logsoftmax_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int64,
                                       C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                       C.c_int32, C.c_void_p]),
    "pcadv_lse_ratio": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int64, C.c_void_p]),
    "pcadv_round_residual": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                                       C.c_void_p, C.c_void_p]),
    "pcadv_bmm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                            C.c_void_p]),
    "pcadv_bmm_tgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                  C.c_void_p]),
    "pcadv_ortho_reg": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pcadv_ortho_reg_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_void_p, C.c_void_p]),
    "pcadv_part_counts": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pcadv_part_iou": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                                 C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pcadv_query_workspace": (C.c_longlong, [C.c_int32, C.c_int64, C.c_int64, C.c_int32]),
    "pcadv_jitter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_ulonglong,
                               C.c_ulonglong, C.c_void_p]),
    "pcadv_bn_stats": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_float, C.c_float,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pcadv_bn_apply": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "pcadv_bn_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                               C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "pcadv_transpose": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "pcadv_version": (C.c_int, []),
    "pcadv_device_check": (C.c_int, []),
    "pcadv_last_error": (C.c_char_p, []),
    "pcadv_launch_count": (C.c_longlong, []),
}

_LIB = None
_DEVICE_OK = False


class PcadvError(RuntimeError):
    pass


def load(build_if_missing=True):
    """dlopen libpcadv.so and bind every declared symbol.  Needs no GPU."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    if not os.path.exists(path):
        if not build_if_missing:
            raise PcadvError("libpcadv.so is not built (%s); run __graft_entry__.build()" % path)
        _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def lib():
    """The loaded library, after checking that the current device is sm_100."""
    global _DEVICE_OK
    l = load()
    if not _DEVICE_OK:
        import torch
        if not torch.cuda.is_available():
            raise PcadvError("libpcadv needs a CUDA device (sm_100a); there is no CPU fallback")
        if l.pcadv_device_check() != 0:
            raise PcadvError(l.pcadv_last_error().decode())
        _DEVICE_OK = True
    return l


def check(rc):
    if rc != 0:
        raise PcadvError("libpcadv call failed (%d): %s" % (rc, load().pcadv_last_error().decode()))


def launch_count():
    return int(load().pcadv_launch_count())
