"""Build the text embedding index — B200 mirror of the reference's scripts/build_text_index.py.

Same inputs and the same on-disk product: a CSV with `text` and `image_path` columns goes in,
`torch.save({"embeddings": (N,d) fp32 unit rows, "image_path": [...], "text": [...]})` comes out
(reference :64-75), readable by both the reference's and this repo's TextSearchIndex.  The
reference's N-iteration batch-1 loop (:57-59) becomes batched clm_encode_text calls.  This script is single
process, like the reference's; the data-parallel builder (one rank per GPU, a shard directory, no collective
on the data path) is scripts/build_image_index.py.
"""
from __future__ import annotations

import argparse
from pathlib import Path
from typing import List, Optional, Sequence

import torch

from ..models.clip_model import load_clip_model


def encode_texts_batched(texts: Sequence[str], model, processor, batch_size: int = 1024) -> torch.Tensor:
    """All captions -> (N, d) fp32 unit rows on the CPU."""
    chunks = []
    for i in range(0, len(texts), batch_size):
        enc = processor(text=list(texts[i:i + batch_size]), return_tensors="pt", padding=True, truncation=True)
        ids = torch.where(enc["attention_mask"].bool(), enc["input_ids"],
                          torch.full_like(enc["input_ids"], model.arch.eos_id))
        chunks.append(model.encode_texts(ids, normalize=True).cpu())
        done = min(i + batch_size, len(texts))
        if done % (batch_size * 8) == 0 or done == len(texts):
            print(f"[build_index] Encoded {done}/{len(texts)} texts")
    return torch.cat(chunks, dim=0) if chunks else torch.empty((0, model.arch.proj_dim))


def save_index(index_path: Path, embeddings: torch.Tensor, image_paths: List[str], texts: List[str]) -> None:
    index_path.parent.mkdir(parents=True, exist_ok=True)
    torch.save({"embeddings": embeddings.float().cpu(), "image_path": list(image_paths),
                "text": list(texts)}, index_path)


def build_text_index(data_csv: Path, index_path: Path, clip_config: Path,
                     lora_dir: Optional[Path] = None, batch_size: int = 1024) -> torch.Tensor:
    import pandas as pd

    print(f"[build_index] Using data CSV : {data_csv}")
    print(f"[build_index] Using LoRA dir: {lora_dir}")
    if not Path(data_csv).exists():
        raise FileNotFoundError(f"CSV not found: {data_csv}")
    df = pd.read_csv(data_csv)
    if "text" not in df.columns or "image_path" not in df.columns:
        raise ValueError("CSV must contain 'text' and 'image_path' columns.")
    if len(df) == 0:
        raise ValueError("CSV is empty.")
    print(f"[build_index] Number of rows: {len(df)}")
    model, processor, device = load_clip_model(config_path=clip_config, use_lora=lora_dir is not None,
                                               lora_weights_path=lora_dir)
    print(f"[build_index] Model loaded on device: {device}")
    texts = df["text"].astype(str).tolist()
    image_paths = df["image_path"].astype(str).tolist()
    emb = encode_texts_batched(texts, model, processor, batch_size)
    # the reference renormalises after stacking (:67); rows are already unit length (clm_l2norm)
    save_index(Path(index_path), emb, image_paths, texts)
    print(f"[build_index] Saved index to: {index_path}")
    print(f"[build_index] Embedding shape: {tuple(emb.shape)}")
    return emb


def main(argv=None):
    root = Path(__file__).resolve().parents[1]
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--data-csv", type=Path, default=Path("data/text/train_fashion.csv"))
    ap.add_argument("--lora-dir", type=Path, default=Path("models/saved/clip-lora/epoch_1"))
    ap.add_argument("--clip-config", type=Path, default=root / "config" / "clip_config.yaml")
    ap.add_argument("--index-path", type=Path, default=Path("data/index/fashion_text_index.pt"))
    ap.add_argument("--batch-size", type=int, default=1024)
    a = ap.parse_args(argv)
    lora_dir = a.lora_dir if a.lora_dir.exists() else None
    build_text_index(a.data_csv, a.index_path, a.clip_config, lora_dir, a.batch_size)


if __name__ == "__main__":
    main()
