"""ctypes binding of libclm_b200.so — the C-ABI declared in include/clm_b200.h.

There is deliberately NO fallback: if the shared library is missing or a call fails, the
caller gets an exception.  torch is used only to hold device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libclm_b200.so"

EPI_NONE = 0
EPI_QUICKGELU = 1
EPI_SPLIT_K = 2
OUT_BF16 = 0
OUT_F32 = 1


class ClmError(RuntimeError):
    pass


class TowerConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("width", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32),
        ("mlp", C.c_int32), ("proj_dim", C.c_int32), ("tokens", C.c_int32), ("image", C.c_int32),
        ("patch", C.c_int32), ("vocab", C.c_int32), ("eos_id", C.c_int32),
        ("lora_cols_qkv", C.c_int32), ("lora_cols_out", C.c_int32), ("lora_cols_fc1", C.c_int32),
        ("lora_cols_fc2", C.c_int32), ("ln_eps", C.c_float),
    ]


_LAYER_FIELDS = [
    "ln1_g", "ln1_b", "w_qkv", "b_qkv", "lora_a_qkv", "lora_b_qkv", "w_o", "b_o", "lora_a_o",
    "lora_b_o", "ln2_g", "ln2_b", "w_fc1", "b_fc1", "w_fc2", "b_fc2", "lora_a_fc1", "lora_b_fc1", "lora_a_fc2",
    "lora_b_fc2",
]
_TOWER_FIELDS = [
    "patch_w", "class_emb", "pre_ln_g", "pre_ln_b", "tok_emb", "pos_emb", "final_ln_g",
    "final_ln_b", "proj_w",
]


class ImageDesc(C.Structure):
    _fields_ = [("data", C.c_void_p), ("height", C.c_int32), ("width", C.c_int32), ("row_stride_bytes", C.c_int32)]


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _LAYER_FIELDS]


_FOLD_FIELDS = [
    "w_qkv_g", "s_qkv", "b_qkv_f", "lora_a_qkv_g", "s_a_qkv",
    "w_fc1_g", "s_fc1", "b_fc1_f", "lora_a_fc1_g", "s_a_fc1",
]


class LayerLnFold(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _FOLD_FIELDS]


class TowerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _TOWER_FIELDS]


_P = C.c_void_p
_I = C.c_int
_F = C.c_float

# name -> (restype, argtypes); mirrors include/clm_b200.h one to one
SIGNATURES = {
    "clm_last_error": (C.c_char_p, []),
    "clm_version": (_I, []),
    "clm_device_check": (_I, []),
    "clm_layernorm": (_I, [_P, _P, _P, _P, _I, _I, _F, _P]),
    "clm_layernorm_ex": (_I, [_P, _I, _P, _P, _P, _I, _I, _F, _P]),
    "clm_embed_text_ex": (_I, [_P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "clm_vision_embed_ln_ex": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "clm_pool_ln_ex": (_I, [_P, _I, _P, _P, _P, _P, _I, _I, _I, _F, _P]),
    "clm_row_stats": (_I, [_P, _P, _I, _I, _F, _P]),
    "clm_gemm_ln_epi": (_I, [_P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P, _I, _I, _P, _P, _P, _I, _I, _P]),
    "clm_fuse_normalize": (_I, [_P, _F, _P, _F, _P, _P, _I, _I, _P]),
    "clm_preprocess_workspace_bytes": (C.c_size_t, [_P, _I, _I]),
    "clm_preprocess_images": (_I, [_P, _I, _I, _P, _P, _P, _P, C.c_size_t, _P]),
    "clm_l2norm": (_I, [_P, _P, _P, _I, _I, _P]),
    "clm_embed_text": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "clm_patch_im2col": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "clm_vision_embed_ln": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P]),
    "clm_pool_ln": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P]),
    "clm_gemm_epi": (_I, [_P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P, _I, _I, _P, _P, _I,
                          _I, _P]),
    "clm_last_gemm_variant": (_I, []),
    "clm_attention": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "clm_quickgelu_fwd": (_I, [_P, _P, C.c_longlong, _P]),
    "clm_quickgelu_bwd": (_I, [_P, _P, _P, C.c_longlong, _P]),
    "clm_layernorm_bwd": (_I, [_P, _I, _P, _P, _P, _P, _I, _I, _F, _I, _P, _I, _I, _P]),
    "clm_transpose_to_bf16": (_I, [_P, _I, C.c_longlong, C.c_longlong, _I, _I, _P, C.c_longlong, C.c_longlong, _I,
                                   _F, _P]),
    "clm_cast_to_bf16": (_I, [_P, _P, C.c_longlong, _F, _P]),
    "clm_lora_wgrad_small": (_I, [_P, _I, _I, _P, _I, _P, _I, _I, _P, _I, _I, _I, _P, _P, _I, _P]),
    "clm_attention_bwd_scratch_bytes": (C.c_size_t, [_I, _I, _I]),
    "clm_attention_bwd": (_I, [_P, _P, _P, _P, C.c_size_t, _I, _I, _I, _I, _P]),
    "clm_clip_loss_workspace_bytes": (C.c_size_t, [_I, _I]),
    "clm_clip_loss": (_I, [_P, _P, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "clm_adamw_step": (_I, [_P, _P, _P, _P, _P, C.c_longlong, _P, _P, _F, _F, _F, _F, _F, _P]),
    "clm_tower_create": (_I, [C.POINTER(TowerConfig), C.POINTER(TowerWeights),
                              C.POINTER(LayerWeights), C.POINTER(_P)]),
    "clm_tower_destroy": (None, [_P]),
    "clm_tower_workspace_bytes": (C.c_size_t, [_P, _I]),
    "clm_tower_set_residual_dtype": (_I, [_P, _I]),
    "clm_tower_residual_dtype": (_I, [_P]),
    "clm_tower_set_ln_fold": (_I, [_P, C.POINTER(LayerLnFold)]),
    "clm_encode_image": (_I, [_P, _P, _I, _P, _I, _P, C.c_size_t, _P]),
    "clm_encode_text": (_I, [_P, _P, _I, _P, _I, _P, C.c_size_t, _P]),
    "clm_encode_text_len": (_I, [_P, _P, _I, _I, _P, _I, _P, C.c_size_t, _P]),
    "clm_search_num_splits": (_I, [_I, _I]),
    "clm_search_topk": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _F, _P, _P, _P]),
    "clm_kth_largest": (_I, [_P, _I, _I, _I, _F, _P, _P]),
    "clm_kth_lower_bound": (_I, [_P, _I, _I, _I, _F, _P, _P]),
    "clm_topk_merge": (_I, [_P, _P, _I, _I, _I, _F, _P, _P, _I, _I, C.c_int64, _P, _P, _P, _P]),
    "clm_topk_gather_chunk_bytes": (C.c_size_t, [_I, _I]),
    "clm_topk_merge_gathered": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "clm_topk_row": (_I, [_P, _I, _I, C.c_int64, _P, _P, _P]),
    "clm_topk_merge_sorted": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "clm_cosine_gemv": (_I, [_P, _P, _I, _I, _P, _P]),
    "clm_launch_count": (C.c_ulonglong, []),
    "clm_launch_count_add": (None, [C.c_longlong]),
    "clm_prof_is_enabled": (_I, []),
    "clm_prof_enable": (_I, [_I]),
    "clm_prof_records": (_I, [C.POINTER(C.c_double), _I]),
    "clm_prof_timeline": (_I, [C.POINTER(C.c_double), _I]),
    "clm_prof_summary": (_I, [_I, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                              C.POINTER(C.c_int)]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises ClmError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    lib_path = Path(os.environ["CLM_LIB_PATH"]) if os.environ.get("CLM_LIB_PATH") else LIB_PATH  # A/B runs of two builds
    if not lib_path.exists():
        raise ClmError(
            f"{lib_path} is missing: build it with `python -m clip_lora_match_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(str(lib_path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().clm_last_error().decode("utf-8", "replace")
        raise ClmError(f"{what or 'clm call'} failed (rc={rc}): {msg}")


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def cur_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


KERNEL_KINDS = {"gemm": 0, "attention": 1, "elementwise": 2, "search": 3, "merge": 4}


def prof_summary(kind: str) -> dict:
    """{ms, flops, bytes, launches} of the launches of one kind recorded since prof_enable(1)."""
    ms, fl, by, n = C.c_double(), C.c_double(), C.c_double(), C.c_int()
    check(load().clm_prof_summary(KERNEL_KINDS[kind], C.byref(ms), C.byref(fl), C.byref(by), C.byref(n)),
          "clm_prof_summary")
    return {"ms": ms.value, "flops": fl.value, "bytes": by.value, "launches": n.value}


def prof_records() -> list:
    """[(kind_name, flops, bytes, ms)] for every launch recorded since prof_enable(1), in order."""
    lib = load()
    n = lib.clm_prof_records(None, 0)
    if n <= 0:
        return []
    buf = (C.c_double * (4 * n))()
    lib.clm_prof_records(buf, n)
    names = {v: k for k, v in KERNEL_KINDS.items()}
    return [(names[int(buf[4 * i])], buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3]) for i in range(n)]


def prof_timeline() -> list:
    """[(kind_name, start_ms, dur_ms)] of every launch recorded since prof_enable(1); start_ms counts from the
    first recorded launch's start event."""
    lib = load()
    n = lib.clm_prof_timeline(None, 0)
    if n <= 0:
        return []
    buf = (C.c_double * (3 * n))()
    lib.clm_prof_timeline(buf, n)
    names = {v: k for k, v in KERNEL_KINDS.items()}
    return [(names[int(buf[3 * i])], buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)]
