"""Report-side service — B200 mirror of the hot-path part of the reference's
src/embedding/finder_service.py.

Kept from the reference: `FinderConfig` fields, `FinderService.report_item(src_image_path, description,
location, reporter, found_at)` with the same rules — `FileNotFoundError` for a missing image (:126-127), the
image copied into `upload_dir` under its own name (:132-136), the stored path relative to `root_dir` with
forward slashes (:138,181), the caption "<description>, ditemukan di <location>" (:158-161), and the quirk
that only the TEXT is embedded: the uploaded image is stored but never encoded (:163; SURVEY.md Appendix D
quirk 5) — and the same result dict (:205-212).

Changed on purpose: the reference appends by load -> torch.cat -> save of the WHOLE index file per reported
item (:172-185, O(N) bytes per item).  Here `index_path` may be a sharded index DIRECTORY
(index_store.ShardedIndexWriter: one new shard per report, O(new rows)); a `.pt` path keeps the reference's
one-file behaviour and its plural metadata keys byte for byte in layout.  YOLO cropping and the Postgres
insert are out of scope: `crop_fn(path) -> path` and `on_item(record) -> id`
hooks take their place.
"""
from __future__ import annotations

import shutil
from dataclasses import dataclass
from datetime import datetime
from pathlib import Path
from typing import Any, Callable, Dict, Optional

import torch

from . import index_store as IS


@dataclass
class FinderConfig:
    """Same fields as the reference (:21-38); `index_path` may also be a sharded index directory."""
    root_dir: Path
    clip_config_path: Path
    lora_dir: Path
    index_path: Path
    upload_dir: Path
    yolo_config_path: Optional[Path] = None


class FinderService:
    def __init__(self, config: FinderConfig, *, model=None, processor=None, device=None,
                 encode_fn: Optional[Callable[[str], torch.Tensor]] = None,
                 crop_fn: Optional[Callable[[Path], Path]] = None,
                 on_item: Optional[Callable[[Dict[str, Any]], Any]] = None) -> None:
        self.config = config
        Path(config.upload_dir).mkdir(parents=True, exist_ok=True)
        if encode_fn is None:
            from ...models.clip_model import encode_text, load_clip_model

            if model is None:
                model, processor, device = load_clip_model(config_path=config.clip_config_path, use_lora=True,
                                                           lora_weights_path=config.lora_dir)
            encode_fn = lambda text: encode_text(text, model, processor, device)  # noqa: E731
        self.model, self.processor, self.device = model, processor, device
        self._encode = encode_fn
        self.crop_fn = crop_fn
        self.on_item = on_item
        self._next_id = 1
        self._writer = None

    # ---- index update ------------------------------------------------------------------------
    def _is_directory_index(self) -> bool:
        p = Path(self.config.index_path)
        return p.is_dir() or p.suffix == ""

    def _append(self, emb: torch.Tensor, image_path: str, text: str) -> None:
        p = Path(self.config.index_path)
        if self._is_directory_index():
            # one writer per service (its constructor scans the directory once) and an incremental manifest
            # update: a report costs O(new rows), not a rescan of every earlier report's sidecar
            if self._writer is None or self._writer.dir != p or self._writer.dim != emb.shape[-1]:
                self._writer = IS.ShardedIndexWriter(p, emb.shape[-1])
            self._writer.append(emb, [image_path], [text])
            IS.append_to_manifest(p, self._writer.last_shard, self._writer.dim)
            return
        # one-file index: the reference's load -> cat -> save (:74-102,172-185), plural keys
        if p.exists():
            data = torch.load(p, map_location="cpu")
            old, paths, texts = data.get("embeddings"), list(data.get("image_paths", [])), list(data.get("texts", []))
        else:
            old, paths, texts = None, [], []
        new = emb if old is None else torch.cat([old, emb], dim=0)
        paths.append(image_path)
        texts.append(text)
        p.parent.mkdir(parents=True, exist_ok=True)
        torch.save({"embeddings": new, "image_paths": paths, "texts": texts}, p)

    # ---- the reference's API -----------------------------------------------------------------
    def report_item(self, src_image_path: Path, description: str, location: Optional[str] = None,
                    reporter: Optional[str] = None, found_at: Optional[datetime] = None) -> Dict[str, Any]:
        src_image_path = Path(src_image_path).resolve()
        if not src_image_path.exists():
            raise FileNotFoundError(f"Source image not found: {src_image_path}")
        if found_at is None:
            found_at = datetime.now()
        dest_path = (Path(self.config.upload_dir) / src_image_path.name).resolve()
        if src_image_path != dest_path:
            shutil.copy2(src_image_path, dest_path)
        rel_image_path = str(dest_path.relative_to(Path(self.config.root_dir).resolve())).replace("\\", "/")
        if self.crop_fn is not None:  # the crop is made for parity with the reference's flow; it is not embedded
            try:
                self.crop_fn(dest_path)
            except Exception as e:  # the reference falls back to the original image (:154-155)
                print(f"[FinderService] crop error, fallback ke gambar asli: {e}")
        full_text = f"{description}, ditemukan di {location}" if location else description
        emb = self._encode(full_text).detach().to("cpu", torch.float32)
        if emb.dim() == 1:
            emb = emb.unsqueeze(0)
        emb = emb / emb.norm(dim=-1, keepdim=True)
        self._append(emb, rel_image_path, full_text)
        record = {"image_path": rel_image_path, "description": full_text, "location": location,
                  "found_at": found_at.isoformat() if found_at else None, "reporter": reporter}
        if self.on_item is not None:
            item_id = self.on_item(dict(record))
        else:
            item_id, self._next_id = self._next_id, self._next_id + 1
        result = {"id": item_id, **record}
        print(f"[FinderService] New item reported: {result}")
        return result
