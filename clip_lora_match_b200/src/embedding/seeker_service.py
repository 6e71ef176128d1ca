"""Seeker-side query fusion and a PERSISTENT index service — B200 mirror of the hot-path part of
the reference's src/embedding/seeker_service.py.

Kept from the reference: `SeekerConfig` fields, `SeekerService.search_items(query_text,
query_image_path, top_k)` and `_build_query_embedding(..., w_text=0.5, w_image=0.5)` with the same
rules (:98-102 ValueError when neither modality is given; :148-151 a single modality is just
renormalised; :153-157 two modalities -> weighted sum -> renormalise) and the same
`FileNotFoundError` for a missing query image (:176-177).

Changed on purpose (Appendix D quirk 6): the reference reloads the index from disk on EVERY query
(:183); here it is loaded once and stays resident on the GPU(s) as fp32 master + bf16 shadow,
which is what makes a 10M-row index searchable at all.  YOLO cropping is out of scope: a
`crop_fn(path) -> path` hook is accepted instead.

Batched form (extension): `fuse_queries(text_embs [Q,d], image_embs [Q,d])` -> [Q,d] on the GPU
through clm_fuse_normalize, then `index.search_batch`.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Callable, List, Optional, Tuple, Union

import torch

from ... import kernels as K
from ...models.clip_model import encode_image, encode_text, load_clip_model
from .search import SearchResult, TextSearchIndex


@dataclass
class SeekerConfig:
    """Same fields as the reference (:19-32); `index_path` may also be a sharded index directory."""
    root_dir: Path
    clip_config_path: Path
    lora_dir: Path
    index_path: Path
    yolo_config_path: Optional[Path] = None
    yolo_crop_dir: Optional[Path] = None


def fuse_queries(text_embs: Optional[torch.Tensor], image_embs: Optional[torch.Tensor],
                 w_text: float = 0.5, w_image: float = 0.5,
                 device: Union[str, torch.device] = "cuda") -> torch.Tensor:
    """[Q,d] (or (d,)) text and/or image embeddings -> fused unit query rows [Q,d] on the GPU."""
    if text_embs is None and image_embs is None:
        raise ValueError("Minimal harus ada query_text atau query_image_path.")

    def prep(t):
        t = t.to(device=device, dtype=torch.float32)
        return (t.unsqueeze(0) if t.dim() == 1 else t).contiguous()

    if text_embs is None or image_embs is None:
        one = prep(text_embs if text_embs is not None else image_embs)
        return K.fuse_normalize(one, 1.0)                      # reference :148-151
    a, b = prep(text_embs), prep(image_embs)
    if a.shape != b.shape:
        raise ValueError(f"text embeddings {tuple(a.shape)} and image embeddings {tuple(b.shape)} differ in shape")
    return K.fuse_normalize(a, w_text, b, w_image)             # reference :153-157


class SeekerService:
    def __init__(self, config: Optional[SeekerConfig] = None, *, model=None, processor=None, device=None,
                 index: Optional[TextSearchIndex] = None,
                 crop_fn: Optional[Callable[[Path], Path]] = None, distributed: bool = False) -> None:
        self.config = config
        if model is None:
            if config is None:
                raise ValueError("either a SeekerConfig or (model, processor, device, index) is required")
            model, processor, device = load_clip_model(config_path=config.clip_config_path, use_lora=True,
                                                       lora_weights_path=config.lora_dir)
        self.model, self.processor, self.device = model, processor, device
        if index is None:
            p = Path(config.index_path)
            index = (TextSearchIndex.from_directory(p, device=device, distributed=distributed) if p.is_dir()
                     else TextSearchIndex(p, device=device, distributed=distributed))
        self.index = index          # resident: NOT reloaded per query (reference :183 does) ...
        self._stamp = self._index_stamp()  # ... but refreshed when the file / manifest on disk has changed
        self.crop_fn = crop_fn

    def _index_stamp(self):
        """(mtime_ns, size) of the index file, or of manifest.json for a shard directory; None without a config."""
        if self.config is None:
            return None
        p = Path(self.config.index_path)
        f = p / "manifest.json" if p.is_dir() else p
        try:
            st = f.stat()
            return (st.st_mtime_ns, st.st_size)
        except OSError:
            return None

    def refresh_if_stale(self) -> bool:
        """Reload the resident index if the finder side has published new rows since it was loaded (one stat()
        per query instead of the reference's torch.load per query, seeker_service.py:183)."""
        stamp = self._index_stamp()
        if stamp is None or stamp == self._stamp:
            return False
        self.reload_index()
        return True

    def reload_index(self) -> None:
        """Explicit refresh after the finder side appended items (replaces the per-query reload)."""
        if self.config is None:
            raise ValueError("reload_index needs a SeekerConfig")
        p = Path(self.config.index_path)
        self.index = (TextSearchIndex.from_directory(p, device=self.device, distributed=self.index.distributed)
                      if p.is_dir() else TextSearchIndex(p, device=self.device, distributed=self.index.distributed))
        self._stamp = self._index_stamp()

    def _build_query_embedding(self, query_text: Optional[str], query_image_path: Optional[Path],
                               w_text: float = 0.5, w_image: float = 0.5) -> torch.Tensor:
        have_text = query_text is not None and query_text.strip() != ""
        have_image = query_image_path is not None
        if not have_text and not have_image:
            raise ValueError("Minimal harus ada query_text atau query_image_path.")
        txt = encode_text(query_text, self.model, self.processor, self.device) if have_text else None
        img = None
        if have_image:
            path = Path(query_image_path)
            if self.crop_fn is not None:
                try:
                    path = Path(self.crop_fn(path))
                except Exception as e:  # the reference falls back to the original image (:136-137)
                    print(f"[SeekerService] crop error, fallback ke gambar asli: {e}")
            img = encode_image(path, self.model, self.processor, self.device)
        return fuse_queries(txt, img, w_text, w_image, self.device)[0].cpu()

    def search_items(self, query_text: Optional[str] = None, query_image_path: Optional[str] = None,
                     top_k: int = 5) -> List[SearchResult]:
        img_path: Optional[Path] = None
        if query_image_path:
            root = self.config.root_dir if self.config is not None else Path(".")
            img_path = (Path(root) / query_image_path).resolve()
            if not img_path.exists():
                raise FileNotFoundError(f"Query image not found: {img_path}")
        query_emb = self._build_query_embedding(query_text=query_text, query_image_path=img_path)
        self.refresh_if_stale()  # items the finder reported since the index was loaded (reference: reload per query)
        return self.index.search_with_embedding(query_emb, top_k=top_k)

    def search_batch(self, text_embs: Optional[torch.Tensor], image_embs: Optional[torch.Tensor],
                     top_k: int = 5, w_text: float = 0.5, w_image: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
        """Config-5 shape: a batch of seeker queries (text and/or image embeddings) against the
        resident finder index -> (scores [Q,k], global ids [Q,k])."""
        return self.index.search_batch(fuse_queries(text_embs, image_embs, w_text, w_image, self.device),
                                       top_k=top_k)
