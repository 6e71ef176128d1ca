"""Sharded, append-only on-disk embedding index.

The reference keeps ONE `.pt` dict and appends by load -> torch.cat -> save of the whole file
(src/embedding/finder_service.py:74-102,172-185: O(N) bytes rewritten per reported item, O(N^2)
for a build).  Here an index is a DIRECTORY of shards:

    <dir>/manifest.json            {"format": "clm-sharded-index/1", "dim": d, "shards": [...]}
    <dir>/shard-000000.pt          {"embeddings": (n_i, d) fp32 unit rows, "image_paths": [...], "texts": [...]}
    <dir>/shard-000001.pt          ...

Every shard file is by itself a valid index of the reference (same dict, the plural key spelling
finder_service.py writes and search.py:41-56 reads), so `TextSearchIndex(shard_path)` of the
reference opens it unchanged; `export_single_file` writes the classic one-file form.

Data-parallel builds need no collective: rank g writes its own shard(s) and a sidecar
(`shard-xxxxxx.json`); the manifest is (re)built by scanning the sidecars, so ranks never have to
agree on anything but the directory.  Appends write a new shard: O(new rows).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

FORMAT = "clm-sharded-index/1"


@dataclass
class ShardInfo:
    file: str
    rows: int
    order: Tuple[int, int]  # (rank or generation, sequence): global row order of the shards

    def to_json(self) -> dict:
        return {"file": self.file, "rows": self.rows, "order": list(self.order)}


def _atomic_write(path: Path, write_fn) -> None:
    tmp = path.with_name(path.name + f".tmp{os.getpid()}")
    write_fn(tmp)
    os.replace(tmp, path)


def _normalize_cpu(e: torch.Tensor) -> torch.Tensor:
    return e / e.norm(dim=-1, keepdim=True)


class ShardedIndexWriter:
    """Writer of one producer (one rank, or the single-process finder).

    `order_major` places this writer's shards in the global row order (rank for a data-parallel
    build; leave 0 for a single writer).  Rows must already be the unit-length embeddings the
    encoder returns (models/clip_model.py:116,148); `renormalize=True` reproduces the reference's
    redundant renormalisation before saving (scripts/build_text_index.py:67)."""

    def __init__(self, directory: Union[str, Path], dim: int, order_major: int = 0, renormalize: bool = False):
        self.dir = Path(directory)
        self.dir.mkdir(parents=True, exist_ok=True)
        self.dim = int(dim)
        self.order_major = int(order_major)
        self.renormalize = renormalize
        existing = [s for s in scan_shards(self.dir) if s.order[0] == self.order_major]
        self._seq = 1 + max((s.order[1] for s in existing), default=-1)

    def append(self, embeddings: torch.Tensor, image_paths: Optional[Sequence[str]] = None,
               texts: Optional[Sequence[str]] = None) -> Path:
        e = embeddings.detach().to("cpu", torch.float32)
        if e.dim() == 1:
            e = e.unsqueeze(0)
        if e.dim() != 2 or e.shape[1] != self.dim:
            raise ValueError(f"embeddings must be (n, {self.dim}), got {tuple(e.shape)}")
        n = e.shape[0]
        image_paths = list(image_paths) if image_paths is not None else [""] * n
        texts = list(texts) if texts is not None else [""] * n
        if len(image_paths) != n or len(texts) != n:
            raise ValueError(f"metadata length mismatch: {n} rows, {len(image_paths)} image paths, {len(texts)} texts")
        if self.renormalize and n:
            e = _normalize_cpu(e)
        # claim the sequence number with an exclusive create, so two producers that share an order_major (two
        # finder processes appending reports) can never write the same shard name
        while True:
            name = f"shard-{self.order_major:03d}-{self._seq:06d}"
            try:
                os.close(os.open(self.dir / (name + ".claim"), os.O_CREAT | os.O_EXCL | os.O_WRONLY))
                break
            except FileExistsError:
                self._seq += 1
        path = self.dir / (name + ".pt")
        _atomic_write(path, lambda p: torch.save({"embeddings": e.contiguous(), "image_paths": image_paths,
                                                  "texts": texts}, p))
        info = ShardInfo(file=path.name, rows=n, order=(self.order_major, self._seq))
        _atomic_write(self.dir / (name + ".json"),
                      lambda p: p.write_text(json.dumps({"dim": self.dim, **info.to_json()})))
        self._seq += 1
        self.last_shard = info
        return path


def scan_shards(directory: Union[str, Path]) -> List[ShardInfo]:
    """Shards of a directory in global row order, from the per-shard sidecars."""
    d = Path(directory)
    out = []
    for side in sorted(d.glob("shard-*.json")):
        try:
            j = json.loads(side.read_text())
        except (OSError, ValueError):
            continue  # a writer is mid-flight; its shard is not published yet
        if (d / j["file"]).exists():
            out.append(ShardInfo(j["file"], int(j["rows"]), (int(j["order"][0]), int(j["order"][1]))))
    out.sort(key=lambda s: s.order)
    return out


def append_to_manifest(directory: Union[str, Path], info: "ShardInfo", dim: int) -> dict:
    """Add ONE new shard to manifest.json without re-reading every sidecar (the finder appends a shard per
    report: rebuilding the manifest each time would be O(reports^2) small-file reads).  Falls back to a full
    rebuild when the manifest is missing, unreadable, of another width, or does not end before the new shard."""
    d = Path(directory)
    try:
        man = json.loads((d / "manifest.json").read_text())
        last = tuple(man["shards"][-1]["order"]) if man["shards"] else (-1, -1)
        if man.get("format") != FORMAT or int(man.get("dim", 0)) not in (0, int(dim)) or last >= tuple(info.order):
            raise ValueError("rebuild")
    except (OSError, ValueError, KeyError, IndexError, TypeError):
        return write_manifest(d)
    man["dim"] = int(dim)
    man["shards"].append(info.to_json())
    man["rows"] = int(man["rows"]) + int(info.rows)
    _atomic_write(d / "manifest.json", lambda p: p.write_text(json.dumps(man, indent=1)))
    return man


def write_manifest(directory: Union[str, Path]) -> dict:
    """(Re)build manifest.json from the sidecars.  Safe to call from any rank, any number of times."""
    d = Path(directory)
    shards = scan_shards(d)
    dims = set()
    for s in shards:
        dims.add(int(json.loads((d / (Path(s.file).stem + ".json")).read_text())["dim"]))
    if len(dims) > 1:
        raise ValueError(f"shards of different embedding widths in {d}: {sorted(dims)}")
    man = {"format": FORMAT, "dim": dims.pop() if dims else 0, "rows": sum(s.rows for s in shards),
           "shards": [s.to_json() for s in shards]}
    _atomic_write(d / "manifest.json", lambda p: p.write_text(json.dumps(man, indent=1)))
    return man


def read_manifest(directory: Union[str, Path]) -> dict:
    d = Path(directory)
    if not d.is_dir():
        raise FileNotFoundError(f"Index directory not found: {d}")
    p = d / "manifest.json"
    if p.exists():
        man = json.loads(p.read_text())
        if man.get("format") != FORMAT:
            raise ValueError(f"{p}: unknown index format {man.get('format')!r}")
        listed = {s["file"] for s in man["shards"]}
        if listed == {s.file for s in scan_shards(d)}:
            return man
    return write_manifest(d)  # missing or stale (shards appended since): rebuild from the sidecars


def shard_row_offsets(man: dict) -> List[int]:
    offs, o = [], 0
    for s in man["shards"]:
        offs.append(o)
        o += int(s["rows"])
    return offs


def load_rows(directory: Union[str, Path], lo: int, hi: int) -> Tuple[torch.Tensor, List[str], List[str]]:
    """Global rows [lo, hi) of the index: only the shard files that intersect the range are read."""
    d = Path(directory)
    man = read_manifest(d)
    dim = int(man["dim"])
    embs, paths, texts = [], [], []
    for s, off in zip(man["shards"], shard_row_offsets(man)):
        n = int(s["rows"])
        a, b = max(lo, off), min(hi, off + n)
        if a >= b:
            continue
        obj = torch.load(d / s["file"], map_location="cpu")
        e = obj["embeddings"].float()
        if e.dim() == 1:
            e = e.unsqueeze(0)
        embs.append(e[a - off:b - off])
        ip = obj.get("image_paths", obj.get("image_path")) or []
        tx = obj.get("texts", obj.get("text")) or []
        paths += [ip[i] if i < len(ip) else "" for i in range(a - off, b - off)]
        texts += [tx[i] if i < len(tx) else "" for i in range(a - off, b - off)]
    e = torch.cat(embs, dim=0) if embs else torch.empty((0, dim), dtype=torch.float32)
    return e, paths, texts


def load_metadata(directory: Union[str, Path]) -> Tuple[List[str], List[str]]:
    """image paths / texts of ALL rows (every rank keeps them: results carry global row ids)."""
    d = Path(directory)
    man = read_manifest(d)
    paths: List[str] = []
    texts: List[str] = []
    for s in man["shards"]:
        obj = torch.load(d / s["file"], map_location="cpu")
        n = int(s["rows"])
        ip = obj.get("image_paths", obj.get("image_path")) or []
        tx = obj.get("texts", obj.get("text")) or []
        paths += [ip[i] if i < len(ip) else "" for i in range(n)]
        texts += [tx[i] if i < len(tx) else "" for i in range(n)]
    return paths, texts


def export_single_file(directory: Union[str, Path], index_path: Union[str, Path], plural_keys: bool = True) -> Dict:
    """The classic one-file index of the reference (`image_paths`/`texts` as finder_service.py:95-102
    writes, or `image_path`/`text` as scripts/build_text_index.py:69-73 does)."""
    man = read_manifest(directory)
    e, paths, texts = load_rows(directory, 0, int(man["rows"]))
    obj = {"embeddings": e}
    obj["image_paths" if plural_keys else "image_path"] = paths
    obj["texts" if plural_keys else "text"] = texts
    index_path = Path(index_path)
    index_path.parent.mkdir(parents=True, exist_ok=True)
    _atomic_write(index_path, lambda p: torch.save(obj, p))
    return obj
