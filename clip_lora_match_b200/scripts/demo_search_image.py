"""Interactive image -> item search — B200 mirror of the reference's scripts/demo_search_image.py
(:13-93): REPL over search_by_image(top_k=3) against the text index."""
from __future__ import annotations

import argparse
from pathlib import Path

from ..models.clip_model import load_clip_model
from ..src.embedding.search import TextSearchIndex


def main(argv=None):
    root = Path(__file__).resolve().parents[1]
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--clip-config", type=Path, default=root / "config" / "clip_config.yaml")
    ap.add_argument("--lora-dir", type=Path, default=Path("models/saved/clip-lora/epoch_1"))
    ap.add_argument("--index-path", type=Path, default=Path("data/index/custom_items_index.pt"))
    ap.add_argument("--top-k", type=int, default=3)
    ap.add_argument("--image", type=Path, default=None, help="run one query image and exit")
    a = ap.parse_args(argv)
    model, processor, device = load_clip_model(config_path=a.clip_config, use_lora=True,
                                               lora_weights_path=a.lora_dir)
    index = TextSearchIndex(a.index_path)
    print("\n=== Image search demo (B200) ===  empty line or 'exit' quits\n")
    while True:
        path = str(a.image) if a.image is not None else input("Image path: ").strip()
        if not path or path.lower() in {"exit", "quit"}:
            break
        try:
            results = index.search_by_image(path, model, processor, device, top_k=a.top_k)
        except FileNotFoundError as e:
            print(f"[demo] {e}")
            if a.image is not None:
                break
            continue
        print(f"\nTop-{len(results)} results for image: {path}")
        for rank, r in enumerate(results, start=1):
            print(f"{rank}. score={r.score:.4f}\n   image: {r.image_path}\n   text : {r.text}")
        print()
        if a.image is not None:
            break


if __name__ == "__main__":
    main()
