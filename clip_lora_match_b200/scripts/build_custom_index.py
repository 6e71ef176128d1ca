"""Build the custom-items text index — B200 mirror of the reference's scripts/build_custom_index.py.

Same input quirk, same product.  The reference reads its CSV with `index_col=0` (reference :33): the file's
header is `image_path,text` but every row holds THREE fields (path, description, location), so pandas makes
the first field the index and the named columns end up shifted — `df.index` is the image path,
`df["image_path"]` the description and `df["text"]` the location (reference :46-53).  The indexed caption is
"<description>, <location>" (reference :56) and the file written is
`{"embeddings": (N,d) fp32 unit rows, "image_path": [...], "text": [...]}` (reference :75-84).
The reference's batch-1 encode loop (:69-72) becomes batched clm_encode_text calls.
"""
from __future__ import annotations

import argparse
from pathlib import Path
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from .build_text_index import encode_texts_batched, save_index


def read_custom_csv(data_csv: Path) -> Tuple[List[str], List[str]]:
    """(image_paths, captions) with the reference's column mapping and its errors (reference :25,33-44)."""
    import pandas as pd

    data_csv = Path(data_csv)
    if not data_csv.exists():
        raise FileNotFoundError(f"CSV not found: {data_csv}")
    df = pd.read_csv(data_csv, header=0, index_col=0)
    if "image_path" not in df.columns or "text" not in df.columns:
        raise ValueError("CSV must contain columns 'image_path' and 'text' (deskripsi & lokasi).")
    if len(df) == 0:
        raise ValueError("CSV is empty.")
    image_paths = df.index.astype(str).tolist()
    desc = df["image_path"].astype(str).tolist()
    loc = df["text"].astype(str).tolist()
    return image_paths, [f"{d}, {l}" for d, l in zip(desc, loc)]


def build_custom_index(data_csv: Path, index_path: Path, encode: Callable[[Sequence[str]], torch.Tensor],
                       log: Callable[[str], None] = print) -> torch.Tensor:
    """CSV -> index file.  `encode` maps a list of captions to (N, d) embeddings (any device); rows are
    re-normalised after stacking as the reference does (:73)."""
    image_paths, texts = read_custom_csv(data_csv)
    log(f"[build_custom_index] Number of rows: {len(texts)}")
    emb = encode(texts).float().cpu()
    emb = emb / emb.norm(dim=-1, keepdim=True)
    save_index(Path(index_path), emb, image_paths, texts)   # keys: embeddings / image_path / text
    log(f"[build_custom_index] Saved index to: {index_path}")
    log(f"[build_custom_index] Embedding shape: {tuple(emb.shape)}")
    return emb


def main(argv=None):
    from ..models.clip_model import load_clip_model

    root = Path(__file__).resolve().parents[1]
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--data-csv", type=Path, default=Path("data/custom/my_items.csv"))
    ap.add_argument("--lora-dir", type=Path, default=Path("models/saved/clip-lora/epoch_1"))
    ap.add_argument("--clip-config", type=Path, default=root / "config" / "clip_config.yaml")
    ap.add_argument("--index-path", type=Path, default=Path("data/index/custom_items_index.pt"))
    ap.add_argument("--batch-size", type=int, default=1024)
    a = ap.parse_args(argv)
    print(f"[build_custom_index] Using data CSV : {a.data_csv}")
    print(f"[build_custom_index] Using LoRA dir: {a.lora_dir}")
    if not a.lora_dir.exists():   # the reference refuses to build without its adapter (:27-28)
        raise FileNotFoundError(f"LoRA checkpoint dir not found: {a.lora_dir}")
    model, processor, device = load_clip_model(config_path=a.clip_config, use_lora=True, lora_weights_path=a.lora_dir)
    print(f"[build_custom_index] Model loaded on device: {device}")
    build_custom_index(a.data_csv, a.index_path, lambda t: encode_texts_batched(t, model, processor, a.batch_size))


if __name__ == "__main__":
    main()
