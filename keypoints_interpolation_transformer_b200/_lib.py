"""ctypes binding of libkit_b200.so (include/kit.h).  The library is the product: if it is
missing or fails to load this module raises -- there is no CPU / PyTorch fallback path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkit_b200.so")

KIT_OK = 0
MASK_NONE, MASK_REPEAT_INC, MASK_KEYPAD_ADD, MASK_TRIANGLE = 0, 1, 2, 4
LOSS_EUCLID, LOSS_MSE, LOSS_DISTANCE = 0, 1, 2
MODEL_COMPLETER, MODEL_CYCLE = 0, 1
MATRIX_TYPES = {"triangle": 0, "repeat": 1, "repeat-inc": 2, "all": 3}
AUG_NONE, AUG_ROTATE, AUG_SHEAR, AUG_ARM_ROTATE = 0, 1, 2, 3
OUT_BF16, OUT_F32, OUT_F32_ATOMIC = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_GELU_BWD = 0, 1, 2


class KitError(RuntimeError):
    pass


class KitModelConfig(C.Structure):
    _fields_ = [("input_size", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32),
                ("heads", C.c_int32), ("ff", C.c_int32), ("max_len", C.c_int32), ("variant", C.c_int32),
                ("reserved", C.c_int32)]


class KitAttnMask(C.Structure):
    _fields_ = [("frame_mask", C.c_void_p), ("frame_mask_stride", C.c_int64), ("flags", C.c_int32),
                ("reserved", C.c_int32), ("bias", C.c_void_p), ("bias_stride_b", C.c_int64),
                ("bias_stride_h", C.c_int64)]


class KitSeqAug(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("cos_t", C.c_float), ("sin_t", C.c_float),
                ("mtx", C.c_double * 9), ("zero_x", C.c_float), ("zero_y", C.c_float),
                ("arm_cos", C.c_float * 8), ("arm_sin", C.c_float * 8)]


class KitPrepassConfig(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("K", C.c_int32), ("normalize", C.c_int32),
                ("left_shoulder", C.c_int32), ("right_shoulder", C.c_int32), ("right_eye", C.c_int32),
                ("n_body", C.c_int32), ("n_hand", C.c_int32), ("arm_chain", C.c_int32 * 8),
                ("zero_masked_enc", C.c_int32), ("k2p", C.c_int32)]


class KitMissingStats(C.Structure):
    _fields_ = [("mean_consecutive_missing", C.c_float), ("std_consecutive_missing", C.c_float),
                ("mean_number_missing_blocks", C.c_float), ("std_number_missing_blocks", C.c_float),
                ("samples", C.c_int32)]


class KitAugPolicy(C.Structure):
    _fields_ = [("prob", C.c_float), ("angle_deg", C.c_float), ("squeeze", C.c_float), ("arm_prob", C.c_float),
                ("has_arms", C.c_int32)]


BUCKET_CALLBACK = C.CFUNCTYPE(None, C.c_int32, C.c_void_p)

_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_SIGNATURES = {
    "kit_last_error": (C.c_char_p, []),
    "kit_version": (C.c_int, []),
    "kit_set_sm_reserve": (C.c_int, [_I32]),
    "kit_layout_num_entries": (_I32, [C.POINTER(KitModelConfig)]),
    "kit_layout_entry": (C.c_int, [C.POINTER(KitModelConfig), _I32, C.c_char_p, _I32, C.POINTER(_I64),
                                   C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I32)]),
    "kit_layout_trainable_floats": (_I64, [C.POINTER(KitModelConfig)]),
    "kit_layout_total_floats": (_I64, [C.POINTER(KitModelConfig)]),
    "kit_layout_num_buckets": (_I32, [C.POINTER(KitModelConfig)]),
    "kit_layout_bucket": (C.c_int, [C.POINTER(KitModelConfig), _I32, C.POINTER(_I64), C.POINTER(_I64)]),
    "kit_engine_create": (C.c_int, [C.POINTER(KitModelConfig), _I32, _I32, _I32, C.POINTER(_P)]),
    "kit_engine_destroy": (C.c_int, [_P]),
    "kit_engine_workspace_bytes": (_I64, [_P]),
    "kit_engine_bind": (C.c_int, [_P, _P, _P, _P, _I64]),
    "kit_engine_refresh_weights": (C.c_int, [_P, _P]),
    "kit_engine_forward": (C.c_int, [_P, _P, _I64, _P, _I64, C.POINTER(KitAttnMask), C.POINTER(KitAttnMask), _I32, _P, _P]),
    "kit_engine_backward": (C.c_int, [_P, _P, _P, _P, _P]),
    "kit_engine_debug_read": (C.c_int, [_P, C.c_char_p, _P, _I64, _P]),
    "kit_engine_last_launches": (_I64, [_P]),
    "kit_engine_set_profiling": (C.c_int, [_P, _I32]),
    "kit_engine_profile_read": (C.c_int, [_P, _I32, C.POINTER(C.c_float), C.POINTER(_I64), C.POINTER(C.c_double)]),
    "kit_prepass": (C.c_int, [C.POINTER(KitPrepassConfig), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "kit_loss_partials": (_I64, [_I64, _I32]),
    "kit_loss_fwd_bwd": (C.c_int, [_P, _P, _P, _I64, _I32, _I32, _F, _P, _P, _P, _P]),
    "kit_cubic_interpolate": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P]),
    "kit_get_mask": (C.c_int, [_P, _I32, _I32, _P, _P]),
    "kit_draw_missing": (C.c_int, [C.POINTER(KitMissingStats), _I32, _I32, C.c_uint64, C.c_uint64, _P, _P, _P, _P, _P]),
    "kit_draw_policy": (C.c_int, [C.POINTER(KitMissingStats), C.POINTER(KitAugPolicy), _I32, _I32, C.c_uint64, _P, _P, _P, _P, _P, _P]),
    "kit_engine_operands": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_I32)]),
    "kit_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _I32, _F, _P]),
    "kit_adam_step_dev": (C.c_int, [_P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _P]),
    "kit_adam_step_dev_range": (C.c_int, [_P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _I32, _P]),
    "kit_gemm_bf16": (C.c_int, [_I32, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _P, _P, _I64, _I32, _I32, _P,
                                _I64, _I32, _P]),
    "kit_ffn_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "kit_ffn_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P]),
    "kit_gemm_lnbwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _P]),
    "kit_attention_fwd": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _I32, _I32, _I32, _I32, _I32,
                                    C.POINTER(KitAttnMask), _P]),
    "kit_attention_bwd": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _P, _I64, _P, _I64, _P, _I64,
                                    _P, _I32, _I32, _I32, _I32, _I32, C.POINTER(KitAttnMask), _P]),
    "kit_umma_probe": (C.c_int, [_P, _I32, _P, _I32, C.POINTER(_I32), _P, _P]),
    "kit_add_layernorm_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "kit_layernorm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "kit_cast_fp32_to_bf16_padded": (C.c_int, [_P, _I64, _I64, _I64, _P, _I64, _P]),
}

_lib = None


def exported_symbols():
    """Every entry point include/kit.h declares."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KitError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != KIT_OK:
        raise KitError(f"kit error {rc}: {lib().kit_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
