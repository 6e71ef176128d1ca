"""Seeker-side search demo — B200 mirror of the reference's scripts/demo_seeker.py (:10-72): text and/or
image query (0.5 / 0.5 fusion) against the finder index, top-3.  The index stays resident on the GPU
between queries (the reference reloads it from disk for every one, seeker_service.py:183)."""
from __future__ import annotations

import argparse
from pathlib import Path

from ..src.embedding.seeker_service import SeekerConfig, SeekerService


def main(argv=None):
    pkg = Path(__file__).resolve().parents[1]
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--root", type=Path, default=Path("."), help="project root the relative image paths refer to")
    ap.add_argument("--clip-config", type=Path, default=pkg / "config" / "clip_config.yaml")
    ap.add_argument("--lora-dir", type=Path, default=Path("models/saved/clip-lora/epoch_1"))
    ap.add_argument("--index-path", type=Path, default=Path("data/index/custom_items_index.pt"),
                    help="one-file .pt index or a sharded index directory")
    ap.add_argument("--top-k", type=int, default=3)
    ap.add_argument("--text", type=str, default=None, help="run one query and exit (no REPL)")
    ap.add_argument("--image", type=str, default=None, help="query image, relative to --root")
    a = ap.parse_args(argv)
    service = SeekerService(SeekerConfig(root_dir=a.root, clip_config_path=a.clip_config, lora_dir=a.lora_dir,
                                         index_path=a.index_path))
    one_shot = a.text is not None or a.image is not None
    print("\n[demo_seeker] text and/or image query; 'exit' quits\n")
    while True:
        q_text = (a.text or "") if one_shot else input("Teks deskripsi (boleh kosong, 'exit' untuk keluar)> ").strip()
        if q_text.lower() in {"exit", "quit"}:
            break
        q_img = a.image if one_shot else (input("Path gambar relatif (boleh kosong)> ").strip() or None)
        if not q_text and q_img is None:
            print("[demo_seeker] Minimal isi teks atau gambar ya.\n")
            if one_shot:
                break
            continue
        try:
            results = service.search_items(query_text=q_text or None, query_image_path=q_img, top_k=a.top_k)
        except FileNotFoundError as e:   # as the reference: report and keep going (:53-58)
            print(f"[demo_seeker] Error: {e}\n")
            results = []
        for i, r in enumerate(results, start=1):
            print(f"{i}. score={r.score:.4f}\n   image: {r.image_path}\n   text : {r.text}")
        print()
        if one_shot:
            break


if __name__ == "__main__":
    main()
