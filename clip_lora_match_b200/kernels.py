"""Thin torch-tensor front end over the C-ABI (one function per entry point).

Every function checks dtypes/contiguity, allocates the output with torch (the caller owns it)
and calls straight into libclm_b200.so on the current CUDA stream.  No arithmetic happens in
Python; if the library is missing these raise (see _lib.load).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import EPI_NONE, EPI_QUICKGELU, OUT_BF16, OUT_F32, check, cur_stream, ptr


def _req(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    _req(x, torch.float32, "x"); _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
    rows, dim = x.shape
    y = torch.empty((rows, dim), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().clm_layernorm(ptr(x), ptr(gamma), ptr(beta), ptr(y), rows, dim, eps, cur_stream()),
          "clm_layernorm")
    return y


def l2norm(x: torch.Tensor, want_bf16: bool = False):
    _req(x, torch.float32, "x")
    rows, dim = x.shape
    y = torch.empty_like(x)
    yb = torch.empty((rows, dim), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    check(_lib.load().clm_l2norm(ptr(x), ptr(y), ptr(yb), rows, dim, cur_stream()), "clm_l2norm")
    return (y, yb) if want_bf16 else y


def fuse_normalize(a: torch.Tensor, wa: float = 1.0, b: Optional[torch.Tensor] = None, wb: float = 0.0,
                   want_bf16: bool = False):
    """normalize(wa*a (+ wb*b)) per row — the seeker query fusion (seeker_service.py:146-157)."""
    _req(a, torch.float32, "a")
    if b is not None:
        _req(b, torch.float32, "b")
        if b.shape != a.shape:
            raise ValueError(f"a is {tuple(a.shape)} but b is {tuple(b.shape)}")
    rows, dim = a.shape
    y = torch.empty_like(a)
    yb = torch.empty((rows, dim), dtype=torch.bfloat16, device=a.device) if want_bf16 else None
    check(_lib.load().clm_fuse_normalize(ptr(a), float(wa), ptr(b), float(wb), ptr(y), ptr(yb), rows, dim,
                                         cur_stream()), "clm_fuse_normalize")
    return (y, yb) if want_bf16 else y


def preprocess_images(images, out_size: int = 224,
                      mean=(0.48145466, 0.4578275, 0.40821073), std=(0.26862954, 0.26130258, 0.27577711),
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Decoded RGB images (list of uint8 CUDA tensors [H, W, 3], any sizes) -> pixel_values fp32
    [B, 3, S, S]: Pillow-exact bicubic shortest-edge resize, centre crop, 1/255, (x-mean)/std."""
    import ctypes as C

    lib = _lib.load()
    b = len(images)
    dev = images[0].device if b else torch.device("cuda")
    if out is None:
        out = torch.empty((b, 3, out_size, out_size), dtype=torch.float32, device=dev)
    if b == 0:
        return out
    descs = (_lib.ImageDesc * b)()
    for i, im in enumerate(images):
        if not im.is_cuda or im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or im.stride(2) != 1 \
                or im.stride(1) != 3:
            raise ValueError(f"image {i} must be a uint8 CUDA tensor [H, W, 3] with packed pixels, got "
                             f"{tuple(im.shape)} {im.dtype} strides {im.stride()}")
        descs[i].data, descs[i].height, descs[i].width = im.data_ptr(), im.shape[0], im.shape[1]
        descs[i].row_stride_bytes = im.stride(0)
    need = lib.clm_preprocess_workspace_bytes(descs, b, out_size)
    if need == 0:
        raise ValueError("clm_preprocess_workspace_bytes rejected the image descriptors")
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    mean_a, std_a = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    check(lib.clm_preprocess_images(descs, b, out_size, mean_a, std_a, ptr(out), ptr(ws), need, cur_stream()),
          "clm_preprocess_images")
    return out


def gemm_epi(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
             residual: Optional[torch.Tensor] = None, act: int = EPI_NONE,
             out_dtype: torch.dtype = torch.bfloat16, a2: Optional[torch.Tensor] = None,
             w2: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = act(a @ w.T (+ a2 @ w2.T) + bias) (+ residual);  a [M,K], w [N,K] bf16."""
    _req(a, torch.bfloat16, "a"); _req(w, torch.bfloat16, "w")
    M, K = a.shape
    N, K2w = w.shape
    if K2w != K:
        raise ValueError(f"a is [*, {K}] but w is [*, {K2w}]")
    if bias is not None:
        _req(bias, torch.float32, "bias")
    if residual is not None:
        _req(residual, torch.float32, "residual")
    k2 = 0
    if a2 is not None:
        _req(a2, torch.bfloat16, "a2"); _req(w2, torch.bfloat16, "w2")
        k2 = a2.shape[1]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    od = OUT_F32 if out.dtype == torch.float32 else OUT_BF16
    check(_lib.load().clm_gemm_epi(
        ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K,
        ptr(a2), a2.stride(0) if a2 is not None else 0, ptr(w2), w2.stride(0) if w2 is not None else 0, k2,
        ptr(out), out.stride(0), od, ptr(bias), ptr(residual),
        residual.stride(0) if residual is not None else 0, act, cur_stream()), "clm_gemm_epi")
    return out


def attention(qkv: torch.Tensor, batch: int, tokens: int, heads: int, causal: bool) -> torch.Tensor:
    _req(qkv, torch.bfloat16, "qkv")
    dim = heads * 64
    if tuple(qkv.shape) != (batch * tokens, 3 * dim):
        raise ValueError(f"qkv must be [{batch * tokens}, {3 * dim}], got {tuple(qkv.shape)}")
    out = torch.empty((batch * tokens, dim), dtype=torch.bfloat16, device=qkv.device)
    check(_lib.load().clm_attention(ptr(qkv), ptr(out), batch, tokens, heads, int(causal), cur_stream()),
          "clm_attention")
    return out


SAMPLE_ROWS = 16384  # rows whose exact scores seed the scan thresholds (multiple of 256)


def search_topk(q_f32: torch.Tensor, q_bf16: torch.Tensor, index_bf16: torch.Tensor,
                index_f32: Optional[torch.Tensor], k: int, id_offset: int = 0,
                margin: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k of q @ index.T: bf16 tensor-core scan + fused per-tile top-kc, then fp32 re-score.

    Returns (scores fp32 [Q,k], ids int64 [Q,k]) sorted descending; ids are global
    (id_offset added).  k must be <= rows of the index shard (the caller clamps, as the
    reference does at src/embedding/search.py:98)."""
    _req(q_f32, torch.float32, "q_f32"); _req(q_bf16, torch.bfloat16, "q_bf16")
    _req(index_bf16, torch.bfloat16, "index_bf16")
    if index_f32 is not None:
        _req(index_f32, torch.float32, "index_f32")
    nq, dim = q_bf16.shape
    n = index_bf16.shape[0]
    if not (1 <= k <= 64):
        raise ValueError("k must be in [1, 64]")
    if margin is None:
        margin = 6 if index_f32 is not None else 0
    kc = max(k, min(64, max(k + margin, 16)))  # candidates kept per (query, split)
    lib = _lib.load()
    dev = q_bf16.device
    # Per-query running threshold shared by all work units of the scan (see clm_search_topk).
    # It is seeded from a sample: scores of the queries against the first SAMPLE_ROWS index rows
    # (plain tcgen05 GEMM, same bf16 operands as the scan) -> exact kc-th largest per query.
    if n >= 4 * SAMPLE_ROWS:
        sample_scores = gemm_epi(q_bf16, index_bf16[:SAMPLE_ROWS], out_dtype=torch.float32)
        thr = torch.empty((nq,), dtype=torch.float32, device=dev)
        check(lib.clm_kth_largest(ptr(sample_scores), nq, SAMPLE_ROWS, kc, 1e-6, ptr(thr), cur_stream()),
              "clm_kth_largest")
        del sample_scores
    else:
        thr = torch.full((nq,), float("-inf"), dtype=torch.float32, device=dev)
    splits = lib.clm_search_num_splits(nq, n)
    cand_s = torch.empty((nq, splits, kc), dtype=torch.float32, device=dev)
    cand_i = torch.empty((nq, splits, kc), dtype=torch.int32, device=dev)
    check(lib.clm_search_topk(ptr(q_bf16), ptr(index_bf16), nq, n, dim, kc, splits, ptr(thr),
                              ptr(cand_s), ptr(cand_i), cur_stream()), "clm_search_topk")
    out_s = torch.empty((nq, k), dtype=torch.float32, device=q_bf16.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=q_bf16.device)
    check(lib.clm_topk_merge(ptr(cand_s), ptr(cand_i), nq, splits, kc, ptr(q_f32), ptr(index_f32), dim,
                             k, id_offset, ptr(out_s), ptr(out_i), cur_stream()), "clm_topk_merge")
    return out_s, out_i


def topk_merge_sorted(scores: torch.Tensor, ids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge [Q, lists, k] sorted per-shard results into the global top-k."""
    _req(scores, torch.float32, "scores"); _req(ids, torch.int64, "ids")
    nq, lists, kk = scores.shape
    if kk != k:
        raise ValueError("last dim must equal k")
    out_s = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=scores.device)
    check(_lib.load().clm_topk_merge_sorted(ptr(scores), ptr(ids), nq, lists, k, ptr(out_s), ptr(out_i),
                                            cur_stream()), "clm_topk_merge_sorted")
    return out_s, out_i


def rescore(cand_score: torch.Tensor, cand_id: torch.Tensor, q_f32: torch.Tensor,
            index_f32: torch.Tensor, k: int, id_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact fp32 scores of nominated rows: cand_* [Q, lists, kc] -> top-k (score desc, id asc)."""
    _req(cand_score, torch.float32, "cand_score"); _req(cand_id, torch.int32, "cand_id")
    _req(q_f32, torch.float32, "q_f32"); _req(index_f32, torch.float32, "index_f32")
    nq, lists, kc = cand_score.shape
    out_s = torch.empty((nq, k), dtype=torch.float32, device=q_f32.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=q_f32.device)
    check(_lib.load().clm_topk_merge(ptr(cand_score), ptr(cand_id), nq, lists, kc, ptr(q_f32),
                                     ptr(index_f32), q_f32.shape[1], k, id_offset, ptr(out_s), ptr(out_i),
                                     cur_stream()), "clm_topk_merge")
    return out_s, out_i


def cosine_gemv(q_f32: torch.Tensor, index_f32: torch.Tensor) -> torch.Tensor:
    """Exact fp32 scores of one (already normalised) query against every index row -> (N,)."""
    _req(q_f32, torch.float32, "q_f32"); _req(index_f32, torch.float32, "index_f32")
    n, dim = index_f32.shape
    if q_f32.numel() != dim:
        raise ValueError("query must have exactly `dim` elements")
    out = torch.empty(n, dtype=torch.float32, device=index_f32.device)
    check(_lib.load().clm_cosine_gemv(ptr(q_f32), ptr(index_f32), n, dim, ptr(out), cur_stream()),
          "clm_cosine_gemv")
    return out
