"""ctypes binding of libmerlin_b200.so (the C ABI declared in include/merlin_b200.h).

The product path has no fallback: if the shared library is missing or a call fails, we raise.
`MERLIN_B200_LIB` may point at an alternative build (e.g. the bring-up build with bounded
mbarrier spins, libmerlin_b200_dbg.so).
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "merlin_b200.h")

MERLIN_OK = 0
EPI_STORE, EPI_RESIDUAL, EPI_QUICKGELU, EPI_SWIGLU = 0, 1, 2, 3


class MerlinError(RuntimeError):
    """Raised when a libmerlin_b200 entry point returns a non-zero status."""


_P = c_void_p
_PROTOTYPES = {
    "merlin_gemm_bf16": [_P, _P, c_int64, _P, c_int64, _P, c_int64, _P, c_int64, _P, c_int, c_int, c_int, c_int, c_int],
    "merlin_gemm_bf16_ex": [_P, _P, c_int64, c_int, _P, c_int64, c_int, _P, c_int64, _P, c_int64, _P, c_int, c_int, c_int, c_int, c_int],
    "merlin_gemm_qkv_rope": [_P, _P, c_int64, _P, c_int64, _P, c_int64, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int],
    "merlin_gemm_qkv_rope_kv": [_P, _P, c_int64, _P, c_int64, _P, c_int64, _P, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                _P, c_int64, c_int, c_int, c_int],
    "merlin_flash_attn_fwd": [_P, _P, c_int64, _P, c_int64, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float],
    "merlin_rmsnorm": [_P, _P, c_int64, _P, _P, c_int64, c_int, c_int, c_float],
    "merlin_layernorm": [_P, _P, c_int64, _P, _P, _P, c_int64, c_int, c_int, c_float],
    "merlin_patch_im2col": [_P, _P, _P, c_int, c_int, c_int, c_int],
    "merlin_vit_embed_ln": [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float],
    "merlin_conv3x3_im2col": [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int],
    "merlin_embed_splice": [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, c_int64, c_int64],
    "merlin_rope_qk_inplace": [_P, _P, c_int64, _P, _P, _P, c_int, c_int, c_int],
    "merlin_rope_qk_inplace_ex": [_P, _P, c_int64, _P, _P, _P, c_int, c_int, c_int, c_int],
    "merlin_flash_attn_bwd": [_P, _P, c_int64, _P, _P, c_int64, _P, _P, _P, c_int64, _P, c_int, c_int, c_int, c_int, c_int, c_float],
    "merlin_cross_entropy_fwd": [_P, _P, c_int64, _P, c_int, c_int, c_int, c_int64, _P, _P],
    "merlin_argmax": [_P, _P, c_int64, c_int, c_int, _P],
    "merlin_rmsnorm_bwd": [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_float],
    "merlin_swiglu_fwd": [_P, _P, _P, c_int, c_int],
    "merlin_swiglu_bwd": [_P, _P, _P, _P, c_int, c_int],
    "merlin_cross_entropy_bwd": [_P, _P, c_int64, _P, c_int, c_int, c_int, c_int64, _P, c_float, _P, c_int64],
    "merlin_embed_splice_bwd": [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int64],
    "merlin_colsum": [_P, _P, c_int64, _P, c_int, c_int],
    "merlin_attn_delta": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int64],
    "merlin_grad_sumsq": [_P, _P, c_int64, _P],
    "merlin_adamw_step": [_P, _P, _P, _P, _P, _P, _P, c_int, c_int64, c_float, c_float, c_float, c_float, c_int, c_float, c_float, _P],
    "merlin_f32_to_bf16": [_P, _P, _P, c_int64],
    "merlin_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_float],
    "merlin_quick_gelu": [_P, _P, _P, _P, c_int64],
    "merlin_conv3x3_col2im": [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int],
    "merlin_vit_embed": [_P, _P, _P, _P, _P, c_int, c_int, c_int],
    "merlin_vit_embed_bwd": [_P, _P, _P, _P, _P, c_int, c_int, c_int],
    "merlin_decode_step": [_P, _P, _P, _P, _P, _P, c_int],
    "merlin_decode_step_batched": [_P, _P, _P, _P, _P, _P, c_int64, c_int],
    "merlin_decode_stream_prepare": [_P, _P, _P],
    "merlin_decode_stream_step": [_P, _P, _P, _P, _P, _P],
    "merlin_embed_token": [_P, _P, _P, _P, c_int, c_int64],
    "merlin_image_preprocess": [_P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, c_int, _P],
    "merlin_kv_append": [_P, _P, c_int64, _P, c_int64, c_int, c_int, c_int, c_int, c_int],
    "merlin_record_advance": [_P, _P, _P, _P, _P, c_int],
    "merlin_dp_load_nccl": [c_char_p],
    "merlin_dp_unique_id": [_P],
    "merlin_dp_init": [_P, c_int, c_int, c_int, ctypes.POINTER(c_void_p)],
    "merlin_dp_allreduce_bucket": [_P, _P, _P, c_int64],
    "merlin_dp_allreduce_f32": [_P, _P, _P, c_int64, c_int],
    "merlin_dp_mem_alloc": [ctypes.POINTER(c_void_p), c_int64],
    "merlin_dp_mem_free": [_P],
    "merlin_dp_register": [_P, _P, c_int64, ctypes.POINTER(c_void_p)],
    "merlin_dp_deregister": [_P, _P],
    "merlin_dp_finalize": [_P],
    "merlin_dp_nccl_version": [],
    "merlin_set_sm_budget": [c_int],
    "merlin_sm_count": [],
}


class DecoderWeights(ctypes.Structure):
    """struct merlin_decoder_weights (include/merlin_b200.h)."""
    _fields_ = [("num_layers", c_int), ("hidden", c_int), ("heads", c_int), ("intermediate", c_int),
                ("vocab", c_int), ("rms_eps", c_float),
                ("qkv_w", ctypes.POINTER(c_void_p)), ("o_w", ctypes.POINTER(c_void_p)),
                ("gate_up_w", ctypes.POINTER(c_void_p)), ("down_w", ctypes.POINTER(c_void_p)),
                ("ln1_w", ctypes.POINTER(c_void_p)), ("ln2_w", ctypes.POINTER(c_void_p)),
                ("norm_w", c_void_p), ("lm_head_w", c_void_p),
                ("kv_cache", ctypes.POINTER(c_void_p)), ("kv_capacity", c_int64),
                ("rope_cos", c_void_p), ("rope_sin", c_void_p)]


class BeamState(ctypes.Structure):
    """struct merlin_beam_state."""
    _fields_ = [("num_beams", c_int), ("suffix_capacity", c_int64), ("prefix_kv", ctypes.POINTER(c_void_p)),
                ("suffix_kv", ctypes.POINTER(c_void_p)), ("pos_dev", c_void_p)]


class LayerWeights(ctypes.Structure):
    """struct merlin_layer_weights."""
    _fields_ = [(n, c_void_p) for n in ("qkv_w", "o_w", "gate_up_w", "down_w", "ln1_w", "ln2_w")]


class VitLayerWeights(ctypes.Structure):
    """struct merlin_vit_layer_weights."""
    _fields_ = [(n, c_void_p) for n in ("qkv_w", "qkv_b", "out_w", "out_b", "ln1_w", "ln1_b", "ln2_w", "ln2_b",
                                        "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class VitWeights(ctypes.Structure):
    """struct merlin_vit_weights."""
    _fields_ = [(n, c_int) for n in ("num_layers", "hidden", "heads", "intermediate", "image_size", "patch_size",
                                     "k_patch_padded")] + [("ln_eps", c_float)] + \
               [(n, c_void_p) for n in ("class_embedding", "patch_weight", "position_embedding", "pre_ln_w", "pre_ln_b")] + \
               [("layers", ctypes.POINTER(VitLayerWeights))]


class VitLayerGrads(ctypes.Structure):
    """struct merlin_vit_layer_grads."""
    _fields_ = [(n, c_void_p) for n in ("qkv_w", "out_w", "fc1_w", "fc2_w", "qkv_b_f32", "out_b_f32", "ln1_w_f32",
                                        "ln1_b_f32", "ln2_w_f32", "ln2_b_f32", "fc1_b_f32", "fc2_b_f32")]


class VitGrads(ctypes.Structure):
    """struct merlin_vit_grads."""
    _fields_ = [(n, c_void_p) for n in ("patch_weight", "position_f32", "class_f32", "pre_ln_w_f32", "pre_ln_b_f32")] + \
               [("layers", ctypes.POINTER(VitLayerGrads))]


_PROTOTYPES["merlin_vit_forward"] = [_P, _P, _P, c_int, _P, _P, c_int]
_PROTOTYPES["merlin_vit_forward_train"] = [_P, _P, _P, c_int, _P, _P, c_int]
_PROTOTYPES["merlin_vit_backward"] = [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int]
_PROTOTYPES["merlin_decoder_layer_forward"] = [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int,
                                               c_int, c_float, c_int]
_PROTOTYPES["merlin_vit_layer_forward"] = [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_int]
class LayerGrads(ctypes.Structure):
    """struct merlin_layer_grads."""
    _fields_ = [(n, c_void_p) for n in ("qkv_w", "o_w", "gate_up_w", "down_w", "ln1_w_f32", "ln2_w_f32")]


_PROTOTYPES["merlin_projector_forward"] = [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_int]
_PROTOTYPES["merlin_lmhead_ce_forward"] = [_P, _P, c_int64, _P, _P, c_int64, _P, c_int, c_int, c_int, c_int, c_int64, _P, _P,
                                          c_int]
_PROTOTYPES["merlin_lmhead_ce_backward"] = [_P, _P, c_int64, _P, c_int, c_int, c_int, c_int, c_int64, _P, c_float, _P,
                                           c_int64, _P, _P, c_int, _P, c_int64, c_int]
_PROTOTYPES["merlin_decoder_layer_forward_train"] = [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                                     c_float, c_int]
_PROTOTYPES["merlin_decoder_layer_backward"] = [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                                c_float, c_int, c_int]
_SIZE_FNS = {"merlin_decoder_layer_saved_bytes": [c_int, c_int, c_int, c_int, c_int, c_int],
             "merlin_decoder_layer_backward_workspace_bytes": [c_int, c_int, c_int, c_int, c_int, c_int],"merlin_decode_workspace_bytes": [c_int, c_int, c_int, c_int],
             "merlin_decode_batched_workspace_bytes": [c_int, c_int, c_int, c_int, c_int],
             "merlin_decode_stream_workspace_bytes": [c_int, c_int, c_int, c_int],
             "merlin_image_preprocess_workspace_bytes": [c_int, c_int, c_int],
             "merlin_projector_workspace_bytes": [c_int, c_int, c_int, c_int],
             "merlin_flash_attn_bwd_workspace_bytes": [c_int, c_int, c_int, c_int],
             "merlin_decoder_layer_workspace_bytes": [c_int, c_int, c_int],
             "merlin_vit_layer_workspace_bytes": [c_int, c_int, c_int],
             "merlin_vit_workspace_bytes": [_P, c_int],
             "merlin_vit_saved_bytes": [_P, c_int],
             "merlin_vit_backward_workspace_bytes": [_P, c_int]}


def declared_symbols(header: str = _HEADER) -> list[str]:
    """Every function name declared in include/merlin_b200.h (used by the symbol-export test)."""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(merlin_[a-z0-9_]+)\s*\(", text)))


def library_path() -> str:
    return os.environ.get("MERLIN_B200_LIB", os.path.join(_HERE, "libmerlin_b200.so"))


_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise MerlinError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C merlin_b200/csrc` — there is no CPU/PyTorch fallback for the hot path")
    lib = ctypes.CDLL(path)
    lib.merlin_last_error.restype = c_char_p
    lib.merlin_last_error.argtypes = []
    for name, argtypes in _SIZE_FNS.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_size_t
        fn.argtypes = argtypes
    for name, argtypes in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != MERLIN_OK:
        msg = load().merlin_last_error().decode("utf-8", "replace")
        raise MerlinError(f"{what} failed with status {status}: {msg}")
