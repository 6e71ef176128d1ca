"""Embedding index + brute-force cosine top-k — B200 mirror of the reference's
src/embedding/search.py.

Reference surface kept (names, argument order, exceptions, result type):
  SearchResult                                      reference :14-20
  TextSearchIndex(index_path)                       reference :24-68
    .search_with_embedding(query_emb, top_k=5)      reference :70-115  ((d,) / (1,d) only)
    .search_by_text / .search_by_image              reference :117-151
Extensions:
  .search_batch(queries [Q,d], top_k) -> (scores [Q,k], ids [Q,k]) device tensors
  row sharding over torch.distributed ranks (each GPU holds a contiguous block of rows, local
  fused top-k, ONE all_gather_into_tensor of the packed Q*k (id, score) pairs, merge in place) — SURVEY.md §8(e).

The index lives on the GPU as an fp32 master (exact re-scoring) plus a bf16 shadow (the copy
the tensor-core scan streams).  Loading accepts both metadata key spellings of the
reference's writers (`image_path`/`text` and `image_paths`/`texts`, reference :41-56).
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Tuple, Union

import torch

from ... import kernels as K
from ...models.clip_model import encode_image, encode_text


@dataclass
class SearchResult:
    """One search hit (reference :14-20)."""
    index: int
    score: float
    image_path: str
    text: str


def _dist_info(distributed: bool) -> Tuple[int, int]:
    if not distributed:
        return 0, 1
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("distributed=True needs an initialised torch.distributed process group")
    return dist.get_rank(), dist.get_world_size()


def shard_bounds(num_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of rows owned by `rank` (balanced to within one row)."""
    return (num_rows * rank) // world, (num_rows * (rank + 1)) // world


def gather_shard_topk(scores: Optional[torch.Tensor], ids: Optional[torch.Tensor], nq: int, k: int,
                      device: torch.device, group=None) -> torch.Tensor:
    """The one exchange step of the sharded search: every rank packs its local top-k (padded to k with
    (id -1, score -inf)) into one chunk -- int64 ids, then fp32 scores (kernels.pack_topk_chunk) -- and ONE
    all_gather_into_tensor collects the chunks rank-major.  Returns the gathered uint8 buffer
    [world * chunk_bytes], which clm_topk_merge_gathered reads in place (no re-layout).  Payload per rank is
    Q*k*12 bytes (0.5 MB at Q=4096, k=10; 2.4 MB at k=50), i.e. latency-bound on NVLink."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    mine = K.pack_topk_chunk(scores, ids, nq, k, device)
    out = torch.empty(world * mine.numel(), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out


def gather_overflow_counts(overflow: Optional[torch.Tensor], device: torch.device, group=None) -> List[int]:
    """How many queries of the last local scan every rank has to redo exactly (kernels.search_topk's overflow
    flags; None = none).  A redo on ANY rank repeats the exchange on ALL of them, so every rank has to see the same
    numbers: a 4-byte all_gather_into_tensor that is enqueued without a host round trip, then ONE host read."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    cnt = (overflow.sum(dtype=torch.int32) if overflow is not None
           else torch.zeros((), dtype=torch.int32, device=device)).reshape(1)
    cnts = torch.empty((world,), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    return cnts.tolist()


def unpack_gathered_topk(gathered: torch.Tensor, world: int, nq: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[Q, world, k] scores / ids views of a gathered buffer (host-side checks and tests; the GPU merge reads
    the buffer in place)."""
    chunk = K.topk_gather_chunk_bytes(nq, k)
    pairs = chunk // 12
    per_rank = gathered.view(world, chunk)
    ids = per_rank[:, : pairs * 8].contiguous().view(torch.int64).view(world, pairs)[:, : nq * k].reshape(world, nq, k)
    sc = per_rank[:, pairs * 8:].contiguous().view(torch.float32).view(world, pairs)[:, : nq * k].reshape(world, nq, k)
    return sc.permute(1, 0, 2).contiguous(), ids.permute(1, 0, 2).contiguous()


def allgather_rows(local: torch.Tensor, total_rows: int, group=None) -> torch.Tensor:
    """Query-side data parallelism of the seeker path (BASELINE configs[4]): rank r encodes queries
    shard_bounds(Q, r, world) of a batch and this gathers the [Q, d] embedding matrix on every rank, so the
    encoder work of a query batch is split over the GPUs instead of being replicated.  One all_gather of
    ceil(Q / world) * d floats per rank (1.6 MB at Q = 4096, d = 768 on 8 GPUs); blocks are padded to equal
    size for the collective and trimmed afterwards."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(total_rows, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} must hold rows {lo}..{hi} ({hi - lo}), got {local.shape[0]}")
    per = (total_rows + world - 1) // world
    buf = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: hi - lo] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    rows = []
    for r, part in enumerate(parts):
        rlo, rhi = shard_bounds(total_rows, r, world)
        rows.append(part[: rhi - rlo])
    return torch.cat(rows, dim=0)


class TextSearchIndex:
    def __init__(self, index_path: Optional[Union[str, Path]] = None, *,
                 embeddings: Optional[torch.Tensor] = None, image_paths: Optional[list] = None,
                 texts: Optional[list] = None, device: Union[str, torch.device] = "cuda",
                 distributed: bool = False, row_offset: Optional[int] = None,
                 total_rows: Optional[int] = None, verbose: bool = True):
        """Load a `.pt` index written by scripts/build_*_index.py (or take tensors directly).

        With distributed=True every rank keeps rows shard_bounds(N, rank, world) of a full index
        file; passing `embeddings` that are already the rank's shard requires row_offset (first
        global row) and total_rows."""
        if index_path is not None:
            index_path = Path(index_path)
            if not index_path.exists():
                raise FileNotFoundError(f"Index file not found: {index_path}")
            obj = torch.load(index_path, map_location="cpu")
            embs = obj.get("embeddings")
            if embs is None:
                raise ValueError("Index file does not contain 'embeddings'")
            images = obj.get("image_paths")
            if images is None:
                images = obj.get("image_path")
            txts = obj.get("texts")
            if txts is None:
                txts = obj.get("text")
            image_paths = list(images) if images is not None else []
            texts = list(txts) if txts is not None else []
        else:
            if embeddings is None:
                raise ValueError("either index_path or embeddings is required")
            embs = embeddings
            image_paths = list(image_paths) if image_paths is not None else []
            texts = list(texts) if texts is not None else []
        if embs.dim() == 1:
            embs = embs.unsqueeze(0)

        self.device = torch.device(device)
        self.rank, self.world = _dist_info(distributed)
        self.distributed = distributed
        self.image_paths: List[str] = image_paths
        self.texts: List[str] = texts

        if distributed and row_offset is None:
            n_total = embs.shape[0]
            lo, hi = shard_bounds(n_total, self.rank, self.world)
            embs = embs[lo:hi]
            self.row_offset, self.num_items = lo, n_total
        else:
            self.row_offset = int(row_offset or 0)
            self.num_items = int(total_rows) if total_rows is not None else embs.shape[0]
        self.dim = embs.shape[1]

        if index_path is not None and self.num_items != len(self.image_paths) and verbose:
            print(f"[TextSearchIndex] WARNING: embeddings rows ({self.num_items}) "
                  f"!= len(image_paths) ({len(self.image_paths)})")
        if verbose:
            print(f"[TextSearchIndex] Loaded {self.num_items} items with dim={self.dim}"
                  + (f" (rank {self.rank}/{self.world}: rows {self.row_offset}.."
                     f"{self.row_offset + embs.shape[0]})" if distributed else ""))

        # row-normalise once, on the GPU, in fp32 (reference :68); keep fp32 master + bf16 shadow
        local = embs.to(device=self.device, dtype=torch.float32).contiguous()
        if local.shape[0] > 0:
            self.embeddings, self.embeddings_bf16 = K.l2norm(local, want_bf16=True)
        else:
            self.embeddings, self.embeddings_bf16 = local, local.to(torch.bfloat16)
        self.local_rows = self.embeddings.shape[0]
        self.last_search_stats: dict = {}  # filled by search_batch: overflow_queries, lists, list_cap

    @classmethod
    def from_directory(cls, directory: Union[str, Path], device: Union[str, torch.device] = "cuda",
                       distributed: bool = False, verbose: bool = True) -> "TextSearchIndex":
        """Open a sharded index directory (index_store.py).  With distributed=True every rank reads
        only the shard files that intersect its row block shard_bounds(N, rank, world) — no rank ever
        holds the whole index — and keeps the (small) metadata of all rows for result lookup."""
        from . import index_store as IS

        man = IS.read_manifest(directory)
        n_total = int(man["rows"])
        rank, world = _dist_info(distributed)
        lo, hi = shard_bounds(n_total, rank, world) if distributed else (0, n_total)
        embs, _, _ = IS.load_rows(directory, lo, hi)
        paths, texts = IS.load_metadata(directory)
        if embs.shape[1] == 0 and int(man["dim"]) > 0:
            embs = torch.empty((0, int(man["dim"])), dtype=torch.float32)
        return cls(embeddings=embs, image_paths=paths, texts=texts, device=device, distributed=distributed,
                   row_offset=lo, total_rows=n_total, verbose=verbose)

    # ---- batched search (extension) ------------------------------------------------------
    def search_batch(self, queries: torch.Tensor, top_k: int = 5) -> Tuple[torch.Tensor, torch.Tensor]:
        """queries [Q, d] (any device) -> (scores fp32 [Q,k], global ids int64 [Q,k]) on the GPU,
        sorted descending, k = min(top_k, N) as the reference (:98) -- any top_k: k <= 0 gives empty [Q, 0]
        results (torch.topk with k = 0), k up to 1024 runs the fused scan, larger k (<= 2048, one GPU) the
        exact per-query path.  On a sharded index every rank passes the same queries and gets the same
        merged result."""
        if queries.dim() != 2:
            raise ValueError(f"queries must be [Q, d], got {tuple(queries.shape)}")
        if queries.shape[-1] != self.dim:
            raise ValueError(f"query_emb dim {queries.shape[-1]} != index dim {self.dim}")
        k = min(int(top_k), self.num_items)
        nq = queries.shape[0]
        if k <= 0 or nq == 0:
            k = max(k, 0)
            return (torch.empty((nq, k), dtype=torch.float32, device=self.device),
                    torch.empty((nq, k), dtype=torch.int64, device=self.device))
        sharded = self.distributed and self.world > 1
        if k > (K.FUSED_MAX_K if sharded else K.EXACT_MAX_K):
            raise ValueError(f"top_k={top_k} (k={k}) exceeds what the search kernels support: "
                             f"{K.FUSED_MAX_K} on a sharded index, {K.EXACT_MAX_K} on one GPU")
        q = queries.to(device=self.device, dtype=torch.float32).contiguous()
        qn, qb = K.l2norm(q, want_bf16=True)  # reference :93
        k_local = min(k, self.local_rows)
        s = i = None
        if not sharded:
            s, i = K.search_topk(qn, qb, self.embeddings_bf16, self.embeddings, k_local,
                                 id_offset=self.row_offset, stats=self.last_search_stats)
            return s, i
        # sharded: the exchange and the cross-shard merge are enqueued BEFORE the host reads the scan's overflow
        # flags, so the GPU does not idle through that round trip; a flagged query (rare: a candidate list that
        # filled up inside the exactness window) is redone exactly and the exchange repeated
        overflow = None
        if k_local > 0:
            s, i, overflow = K.search_topk(qn, qb, self.embeddings_bf16, self.embeddings, k_local,
                                           id_offset=self.row_offset, stats=self.last_search_stats,
                                           defer_overflow=True)
        out = K.topk_merge_gathered(gather_shard_topk(s, i, nq, k, self.device), self.world, nq, k)
        counts = gather_overflow_counts(overflow, self.device)  # the ONE host synchronisation of the call
        self.last_search_stats["overflow_queries"] = counts[self.rank]
        if any(counts):
            if counts[self.rank]:
                K.resolve_overflow(overflow, qn, self.embeddings, k_local, self.row_offset, s, i)
            out = K.topk_merge_gathered(gather_shard_topk(s, i, nq, k, self.device), self.world, nq, k)
        return out

    # ---- reference API -------------------------------------------------------------------
    def search_with_embedding(self, query_emb: torch.Tensor, top_k: int = 5) -> List[SearchResult]:
        """Single-query search; shape rules and errors as the reference (:80-90)."""
        if query_emb.ndim == 1:
            query_emb = query_emb.unsqueeze(0)
        elif query_emb.ndim != 2 or query_emb.shape[0] != 1:
            raise ValueError(f"query_emb must be shape (d,) or (1, d), got {tuple(query_emb.shape)}")
        if query_emb.shape[-1] != self.dim:
            raise ValueError(f"query_emb dim {query_emb.shape[-1]} != index dim {self.dim}")
        scores, indices = self.search_batch(query_emb, top_k=top_k)
        results: List[SearchResult] = []
        for idx, score in zip(indices[0].tolist(), scores[0].tolist()):
            img = self.image_paths[idx] if idx < len(self.image_paths) else ""
            txt = self.texts[idx] if idx < len(self.texts) else ""
            results.append(SearchResult(index=idx, score=float(score), image_path=img, text=txt))
        return results

    def search_by_text(self, query: str, model, processor, device: torch.device, top_k: int = 5) -> List[SearchResult]:
        query_emb = encode_text(query, model, processor, device)
        return self.search_with_embedding(query_emb, top_k=top_k)

    def search_by_image(self, image_path: Union[str, Path], model, processor, device: torch.device,
                        top_k: int = 5) -> List[SearchResult]:
        query_emb = encode_image(image_path, model, processor, device)
        return self.search_with_embedding(query_emb, top_k=top_k)
