"""Cosine similarity + top-k — B200 mirror of the reference's src/embedding/similarity.py.

Same names, argument meaning and return types as the reference (cosine_similarity :10-33,
top_k_similar :36-58).  Inputs may live on any device; the arithmetic runs in the sm_100a
kernels (clm_l2norm, clm_cosine_gemv, clm_search_topk, clm_topk_merge) and results come back
on the input's device.  Batched queries [Q, d] are accepted by top_k_similar as an extension.
"""
from __future__ import annotations

from typing import Tuple

import torch

from ... import kernels as K


def _cuda(t: torch.Tensor) -> torch.Tensor:
    return t.to(device="cuda", dtype=torch.float32).contiguous()


def cosine_similarity(query: torch.Tensor, candidates: torch.Tensor) -> torch.Tensor:
    """One query (d,)/(1,d) against candidates (N,d) -> (N,) fp32.  Both sides are re-normalised
    without epsilon (reference :28-29); scores are exact fp32 dot products."""
    if query.dim() == 1:
        query = query.unsqueeze(0)
    if query.shape[0] != 1:
        raise ValueError("cosine_similarity takes a single query; use top_k_similar for batches")
    dev = query.device
    q = K.l2norm(_cuda(query))
    c = K.l2norm(_cuda(candidates))
    return K.cosine_gemv(q, c).to(dev)


def top_k_similar(query: torch.Tensor, candidates: torch.Tensor, k: int = 5) -> Tuple[torch.Tensor, torch.Tensor]:
    """(topk_values (k,), topk_indices (k,) int64) with k = min(k, N) (reference :56-57).
    A [Q, d] query batch returns [Q, k] tensors."""
    single = query.dim() == 1
    if single:
        query = query.unsqueeze(0)
    squeeze = single or query.shape[0] == 1
    dev = query.device
    q, qb = K.l2norm(_cuda(query), want_bf16=True)
    c, cb = K.l2norm(_cuda(candidates), want_bf16=True)
    k = min(k, c.shape[0])
    values, indices = K.search_topk(q, qb, cb, c, k)
    if squeeze:
        values, indices = values[0], indices[0]
    return values.to(dev), indices.to(dev)
