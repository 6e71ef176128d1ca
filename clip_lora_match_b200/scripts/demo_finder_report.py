"""Finder-side report demo — B200 mirror of the reference's scripts/demo_finder_report.py (:10-42): one
`FinderService.report_item` call (image stored, caption embedded, index appended) and its result dict."""
from __future__ import annotations

import argparse
from datetime import datetime
from pathlib import Path

from ..src.embedding.finder_service import FinderConfig, FinderService


def main(argv=None):
    pkg = Path(__file__).resolve().parents[1]
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--root", type=Path, default=Path("."))
    ap.add_argument("--image", type=Path, required=True, help="the found item's photo")
    ap.add_argument("--description", type=str, required=True)
    ap.add_argument("--location", type=str, default=None)
    ap.add_argument("--reporter", type=str, default=None)
    ap.add_argument("--clip-config", type=Path, default=pkg / "config" / "clip_config.yaml")
    ap.add_argument("--lora-dir", type=Path, default=Path("models/saved/clip-lora/epoch_1"))
    ap.add_argument("--index-path", type=Path, default=Path("data/index/custom_items_index.pt"),
                    help="one-file .pt index (reference behaviour) or a directory (one shard per report)")
    ap.add_argument("--upload-dir", type=Path, default=Path("data/reported/images"))
    a = ap.parse_args(argv)
    root = a.root.resolve()
    service = FinderService(FinderConfig(root_dir=root, clip_config_path=a.clip_config, lora_dir=a.lora_dir,
                                         index_path=root / a.index_path, upload_dir=root / a.upload_dir))
    result = service.report_item(src_image_path=a.image, description=a.description, location=a.location,
                                 reporter=a.reporter, found_at=datetime.now())
    print("\n[demo_finder_report] Report selesai:")
    print(result)


if __name__ == "__main__":
    main()
