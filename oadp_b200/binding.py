"""ctypes binding of liboake_b200.so (include/oake_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails this raises.
"""
from __future__ import annotations

import ctypes as C
import pathlib

LIB_PATH = pathlib.Path(__file__).resolve().parent / 'liboake_b200.so'

VARIANT_T50 = 0
VARIANT_T197 = 1
DTYPE_F32, DTYPE_F16, DTYPE_BF16 = 0, 1, 2


class OakeError(RuntimeError):
    pass


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        'qkv_w', 'qkv_s', 'qkv_c', 'out_w', 'out_b', 'fc1_w', 'fc1_s', 'fc1_c', 'fc2_w', 'fc2_b')]


class Weights(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('layers', 'width', 'heads', 'patch', 'out_dim', 'image')] + [
        (n, C.c_void_p) for n in ('conv1_w', 'class_emb', 'pos_t50', 'pos_t197', 'ln_pre_w',
                                  'ln_pre_b', 'ln_post_w', 'ln_post_b', 'proj_w')
    ] + [('layer', C.POINTER(LayerWeights))]


class TextWeights(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('layers', 'width', 'heads', 'vocab', 'context', 'out_dim')] + [
        (n, C.c_void_p) for n in ('token_emb', 'pos', 'ln_final_w', 'ln_final_b', 'proj_w')
    ] + [('layer', C.POINTER(LayerWeights))]


# name -> (restype, argtypes); every symbol include/oake_b200.h declares
SIGNATURES = {
    'oake_create': (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Weights)]),
    'oake_destroy': (None, [C.c_void_p]),
    'oake_workspace_bytes': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    'oake_encode_pixels': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'oake_encode_crops_u8': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'oake_resize_u8': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'oake_resize_to_patches': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                         C.c_void_p, C.c_void_p]),
    'oake_encode_patches': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_size_t, C.c_void_p]),
    'oake_object_masks': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'oake_classifier_workspace_bytes': (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    'oake_normalized_linear_fwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'oake_normalized_linear_bwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                             C.c_void_p]),
    'oake_cosine_logits_fwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                         C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_void_p]),
    'oake_cosine_logits_bwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_size_t, C.c_void_p]),
    'oake_classifier_prepare': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    'oake_classifier_fwd': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_size_t, C.c_void_p]),
    'oake_vild_ensemble': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int, C.c_void_p]),
    'oake_softmax_rows': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    'oake_nms_workspace_bytes': (C.c_int, [C.c_int, C.POINTER(C.c_size_t)]),
    'oake_multiclass_nms': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    'oake_loss_workspace_bytes': (C.c_int, [C.c_int, C.POINTER(C.c_size_t)]),
    'oake_pair_loss': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    'oake_rkd_loss': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_size_t, C.c_void_p]),
    'oake_asymmetric_loss': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.c_void_p]),
    'oake_jpeg_desc_bytes': (C.c_size_t, []),
    'oake_jpeg_parse': (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    'oake_jpeg_stream_bound': (C.c_size_t, [C.c_void_p]),
    'oake_jpeg_stage': (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_uint64,
                                  C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_uint64)]),
    'oake_jpeg_decode': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    'oake_text_create': (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(TextWeights)]),
    'oake_text_destroy': (None, [C.c_void_p]),
    'oake_text_workspace_bytes': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    'oake_encode_text': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                   C.c_void_p]),
    'oake_last_error': (C.c_char_p, []),
    'oake_act_dtype': (C.c_char_p, []),
    'oake_abi_version': (C.c_int, []),
    'oake_launch_count': (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong)]),
    'oake_profile_enable': (C.c_int, [C.c_void_p, C.c_int]),
    'oake_profile_collect': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_int)]),
    'oake_test_gemm': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p]),
    'oake_test_layernorm': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'oake_test_attention_main': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'oake_test_attention_side': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p]),
    'oake_test_im2col': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'oake_test_patch_embed': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise OakeError(
            f'{LIB_PATH} is missing -- build it with `python -m oadp_b200.build` '
            '(nvcc, sm_100a). There is no CPU or PyTorch fallback for the OAKE hot path.')
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.oake_abi_version() != 1:
        raise OakeError(f'ABI version mismatch: library reports {lib.oake_abi_version()}')
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise OakeError(load().oake_last_error().decode() or f'oake call failed rc={rc}')


def act_dtype_name() -> str:
    return load().oake_act_dtype().decode()
