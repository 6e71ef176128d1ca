"""Build an IMAGE embedding index, data-parallel and sharded (BASELINE.json configs[2]; SURVEY.md
§8f rank 1).  The reference only ever indexes captions one at a time (scripts/build_text_index.py:57-59;
Appendix D quirk 5); its unused batched surface is src/embedding/embed_image.py:57-98, which this
drives at full batch on every GPU.

Each rank encodes the contiguous item block shard_bounds(N, rank, world) and appends its embeddings
to its own shards of a sharded index directory (src/embedding/index_store.py): no collective on the
data path, and the rank's shard is what it later scans in TextSearchIndex.from_directory(...,
distributed=True).  `--export` additionally writes the reference's one-file `.pt`.

    python -m clip_lora_match_b200.scripts.build_image_index --csv data/custom/items.csv --out data/index/items
    torchrun --nproc-per-node 8 -m clip_lora_match_b200.scripts.build_image_index --synthetic 1000000 --out /tmp/idx
"""
from __future__ import annotations

import argparse
import os
from pathlib import Path
from typing import Callable, Iterator, List, Optional, Sequence, Tuple

import torch

from ..models.clip_model import load_clip_model
from ..src.embedding import index_store as IS
from ..src.embedding.search import shard_bounds

Batch = Tuple[torch.Tensor, List[str], List[str]]  # pixel_values [b,3,H,W], image paths, texts


def image_file_batches(paths: Sequence[str], texts: Sequence[str], processor, batch_size: int) -> Iterator[Batch]:
    """Decode + CLIP-preprocess image files on the host (models/clip_model.py:105-107), batch_size at a time."""
    from PIL import Image

    for i in range(0, len(paths), batch_size):
        chunk = list(paths[i:i + batch_size])
        for p in chunk:
            if not Path(p).exists():
                raise FileNotFoundError(f"Image not found: {p}")
        imgs = [Image.open(p).convert("RGB") for p in chunk]
        pv = processor(images=imgs, return_tensors="pt")["pixel_values"]
        yield pv, chunk, list(texts[i:i + batch_size])


def synthetic_batches(lo: int, hi: int, image: int, batch_size: int, device, seed: int = 2) -> Iterator[Batch]:
    """Seeded N(0,1) pixel_values generated on the device (BASELINE.json: synthetic 224px images)."""
    for i in range(lo, hi, batch_size):
        n = min(batch_size, hi - i)
        g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + i)
        pv = torch.randn((n, 3, image, image), generator=g, device=device)
        yield pv, [f"synthetic://{j}" for j in range(i, i + n)], [""] * n


def build_image_index(batches: Iterator[Batch], out_dir: Path, model, rank: int = 0,
                      rows_per_shard: int = 262_144, log: Callable[[str], None] = print) -> int:
    """Encode every batch and append the embeddings to this rank's shards; returns rows written."""
    writer = IS.ShardedIndexWriter(out_dir, model.arch.proj_dim, order_major=rank)
    pend_e, pend_p, pend_t, pending, total = [], [], [], 0, 0

    def flush():
        nonlocal pend_e, pend_p, pend_t, pending
        if pending:
            writer.append(torch.cat(pend_e, dim=0), pend_p, pend_t)
        pend_e, pend_p, pend_t, pending = [], [], [], 0

    for pv, paths, texts in batches:
        emb = model.encode_images(pv, normalize=True)  # unit rows (clm_l2norm), on the GPU
        pend_e.append(emb.cpu()); pend_p += paths; pend_t += texts
        pending += emb.shape[0]; total += emb.shape[0]
        if pending >= rows_per_shard:
            flush()
            log(f"[build_image_index] rank {rank}: {total} rows written")
    flush()
    return total


def main(argv=None):
    root = Path(__file__).resolve().parents[1]
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--csv", type=Path, default=None, help="CSV with an image_path column (and optional text)")
    ap.add_argument("--synthetic", type=int, default=0, help="index N synthetic images instead of files")
    ap.add_argument("--out", type=Path, required=True, help="sharded index directory")
    ap.add_argument("--export", type=Path, default=None, help="also write the reference's one-file .pt (rank 0)")
    ap.add_argument("--clip-config", type=Path, default=root / "config" / "clip_config.yaml")
    ap.add_argument("--lora-dir", type=Path, default=Path("models/saved/clip-lora/epoch_1"))
    ap.add_argument("--batch-size", type=int, default=1024)
    a = ap.parse_args(argv)

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    lora_dir = a.lora_dir if a.lora_dir.exists() else None
    model, processor, device = load_clip_model(config_path=a.clip_config, use_lora=lora_dir is not None,
                                               lora_weights_path=lora_dir)
    if a.synthetic > 0:
        lo, hi = shard_bounds(a.synthetic, rank, world)
        batches = synthetic_batches(lo, hi, model.arch.image, a.batch_size, device)
    else:
        import pandas as pd

        if a.csv is None or not a.csv.exists():
            raise FileNotFoundError(f"CSV not found: {a.csv}")
        df = pd.read_csv(a.csv)
        if "image_path" not in df.columns:
            raise ValueError("CSV must contain an 'image_path' column.")
        paths = df["image_path"].astype(str).tolist()
        texts = df["text"].astype(str).tolist() if "text" in df.columns else [""] * len(paths)
        lo, hi = shard_bounds(len(paths), rank, world)
        batches = image_file_batches(paths[lo:hi], texts[lo:hi], processor, a.batch_size)
    n = build_image_index(batches, a.out, model, rank=rank)
    print(f"[build_image_index] rank {rank}/{world}: {n} rows (global rows {lo}..{hi})")
    if world > 1:
        import torch.distributed as dist

        dist.barrier()  # control plane only: every rank's shards are on disk before the manifest is cut
    if rank == 0:
        man = IS.write_manifest(a.out)
        print(f"[build_image_index] manifest: {man['rows']} rows, dim {man['dim']}, {len(man['shards'])} shards")
        if a.export is not None:
            IS.export_single_file(a.out, a.export)
            print(f"[build_image_index] exported {a.export}")
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
