"""Rebuild the index from the stored items — B200 mirror of the reference's scripts/rebuild_index.py.

The reference walks the `found_items` table ordered by id (reference :47-52), encodes every item's
description one at a time (:64-80) and writes `{"embeddings", "image_paths", "texts"}` — the PLURAL key
spelling (:82-96), the one `FinderService._load_index` reads.  The database layer is out of scope here
(SURVEY.md §2 row 19), so the items come in as records — anything with `.id`, `.description`, `.image_path`
attributes or the same dict keys (a SQLAlchemy row qualifies) — or from a JSON-lines dump of the table; the
encode loop becomes batched clm_encode_text calls and the product is the same file.
"""
from __future__ import annotations

import argparse
import json
from pathlib import Path
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch


def _field(item, name: str):
    return item[name] if isinstance(item, dict) else getattr(item, name)


def items_to_rows(items: Iterable) -> Tuple[List[str], List[str]]:
    """(image_paths, texts) in id order, as the reference's `order_by(FoundItem.id)` (:49)."""
    rows = sorted(items, key=lambda it: _field(it, "id"))
    return [str(_field(it, "image_path")) for it in rows], [str(_field(it, "description")) for it in rows]


def read_items_jsonl(path: Path) -> List[dict]:
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"items dump not found: {path}")
    return [json.loads(line) for line in path.read_text().splitlines() if line.strip()]


def rebuild_index(items: Iterable, index_path: Path, encode: Callable[[Sequence[str]], torch.Tensor],
                  log: Callable[[str], None] = print) -> Optional[torch.Tensor]:
    """Items -> index file with the plural metadata keys.  An empty table writes nothing (reference :54-56).
    `encode` maps a list of descriptions to (N, d) embeddings; rows are normalised as the reference does
    per item (:72)."""
    image_paths, texts = items_to_rows(items)
    log(f"Found {len(texts)} items")
    if not texts:
        log("No items. Nothing to rebuild.")
        return None
    emb = encode(texts).float().cpu()
    emb = emb / emb.norm(dim=-1, keepdim=True)
    index_path = Path(index_path)
    index_path.parent.mkdir(parents=True, exist_ok=True)
    torch.save({"embeddings": emb, "image_paths": image_paths, "texts": texts}, index_path)
    data = torch.load(index_path)   # the reference's own verification step (:103-112)
    ok = data["embeddings"].shape[0] == len(data["image_paths"]) == len(data["texts"])
    log(f"Index saved to: {index_path} ({'synchronized' if ok else 'NOT synchronized'})")
    return emb


def main(argv=None):
    from ..models.clip_model import load_clip_model
    from .build_text_index import encode_texts_batched

    root = Path(__file__).resolve().parents[1]
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--items", type=Path, required=True, help="JSON-lines dump of found_items (id, description, image_path)")
    ap.add_argument("--lora-dir", type=Path, default=Path("models/saved/clip-lora/epoch_1"))
    ap.add_argument("--clip-config", type=Path, default=root / "config" / "clip_config.yaml")
    ap.add_argument("--index-path", type=Path, default=Path("data/index/custom_items_index.pt"))
    ap.add_argument("--batch-size", type=int, default=1024)
    a = ap.parse_args(argv)
    model, processor, device = load_clip_model(config_path=a.clip_config, use_lora=a.lora_dir.exists(),
                                               lora_weights_path=a.lora_dir if a.lora_dir.exists() else None)
    print(f"Model loaded on {device}")
    rebuild_index(read_items_jsonl(a.items), a.index_path,
                  lambda t: encode_texts_batched(t, model, processor, a.batch_size))


if __name__ == "__main__":
    main()
