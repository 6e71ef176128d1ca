"""Interactive text -> item search — B200 mirror of the reference's scripts/demo_search_text.py
(:11-56): load CLIP+LoRA, load the index once, REPL over search_by_text(top_k=5)."""
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
    ap.add_argument("--index-path", type=Path, default=Path("data/index/fashion_text_index.pt"))
    ap.add_argument("--top-k", type=int, default=5)
    ap.add_argument("--query", type=str, default=None, help="run one query and exit (no REPL)")
    a = ap.parse_args(argv)
    model, processor, device = load_clip_model(config_path=a.clip_config, use_lora=True,
                                               lora_weights_path=a.lora_dir)
    index = TextSearchIndex(a.index_path)
    print("\n=== Text search demo (B200) ===  empty line or 'exit' quits\n")
    while True:
        query = a.query if a.query is not None else input("Query: ").strip()
        if not query or query.lower() in {"exit", "quit"}:
            break
        results = index.search_by_text(query, model, processor, device, top_k=a.top_k)
        print(f"\nTop-{len(results)} results for: '{query}'")
        for rank, r in enumerate(results, start=1):
            print(f"{rank}. score={r.score:.4f}\n   image: {r.image_path}\n   text : {r.text}")
        print()
        if a.query is not None:
            break


if __name__ == "__main__":
    main()
