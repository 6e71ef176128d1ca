"""CPU tests of the sharded on-disk index: shard files stay readable as
reference-format indices, appends are O(new rows), rank-wise writers need no coordination, and a
row range touches only the shards it intersects."""
import json
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clip_lora_match_b200.src.embedding import index_store as IS
from clip_lora_match_b200.src.embedding.search import shard_bounds
from oracle import clip_oracle as O


def _rows(n, d, seed):
    return O.synth_unit_rows(n, d, seed)


def test_append_scan_manifest_roundtrip(tmp_path):
    w = IS.ShardedIndexWriter(tmp_path / "idx", dim=32)
    a, b = _rows(5, 32, 1), _rows(3, 32, 2)
    p0 = w.append(a, [f"a{i}.jpg" for i in range(5)], [f"ta{i}" for i in range(5)])
    w.append(b, [f"b{i}.jpg" for i in range(3)], None)
    man = IS.read_manifest(tmp_path / "idx")
    assert man["format"] == IS.FORMAT and man["rows"] == 8 and man["dim"] == 32 and len(man["shards"]) == 2
    # every shard is a valid index of the reference: the dict finder_service.py:95-102 writes
    obj = torch.load(p0, map_location="cpu")
    assert set(obj) == {"embeddings", "image_paths", "texts"} and obj["embeddings"].dtype == torch.float32
    e, paths, texts = IS.load_rows(tmp_path / "idx", 0, 8)
    assert torch.equal(e, torch.cat([a, b])) and paths[5] == "b0.jpg" and texts[:5] == [f"ta{i}" for i in range(5)]
    assert texts[5:] == ["", "", ""]
    # a range inside the second shard reads only that shard's rows
    e2, p2, _ = IS.load_rows(tmp_path / "idx", 6, 8)
    assert torch.equal(e2, b[1:]) and p2 == ["b1.jpg", "b2.jpg"]


def test_append_after_manifest_is_picked_up_and_writer_resumes(tmp_path):
    d = tmp_path / "idx"
    IS.ShardedIndexWriter(d, dim=16).append(_rows(4, 16, 1))
    assert IS.read_manifest(d)["rows"] == 4
    w2 = IS.ShardedIndexWriter(d, dim=16)          # a new process appends later (FinderService.report_item)
    w2.append(_rows(1, 16, 2)[0])                  # a single (d,) vector
    man = IS.read_manifest(d)                      # stale manifest is rebuilt from the sidecars
    assert man["rows"] == 5 and [s["order"][1] for s in man["shards"]] == [0, 1]


def test_export_single_file_matches_reference_layouts(tmp_path):
    d = tmp_path / "idx"
    w = IS.ShardedIndexWriter(d, dim=8)
    a = _rows(6, 8, 3)
    w.append(a[:2], ["x", "y"], ["tx", "ty"]); w.append(a[2:], list("abcd"), list("ABCD"))
    obj = IS.export_single_file(d, tmp_path / "one.pt")
    assert set(obj) == {"embeddings", "image_paths", "texts"} and torch.equal(obj["embeddings"], a)
    obj2 = IS.export_single_file(d, tmp_path / "two.pt", plural_keys=False)
    assert set(obj2) == {"embeddings", "image_path", "text"}      # scripts/build_text_index.py:69-73 spelling
    back = torch.load(tmp_path / "two.pt", map_location="cpu")
    assert back["text"] == ["tx", "ty", "A", "B", "C", "D"]


def test_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        IS.read_manifest(tmp_path / "missing")
    w = IS.ShardedIndexWriter(tmp_path / "idx", dim=8)
    with pytest.raises(ValueError):
        w.append(_rows(2, 16, 1))
    with pytest.raises(ValueError):
        w.append(_rows(2, 8, 1), ["only-one"], None)
    IS.ShardedIndexWriter(tmp_path / "idx", dim=8).append(_rows(1, 8, 1))
    IS.ShardedIndexWriter(tmp_path / "idx", dim=4, order_major=1).append(_rows(1, 4, 1))
    with pytest.raises(ValueError):
        IS.write_manifest(tmp_path / "idx")        # mixed widths


def test_empty_directory_and_empty_ranges(tmp_path):
    (tmp_path / "idx").mkdir()
    man = IS.read_manifest(tmp_path / "idx")
    assert man["rows"] == 0 and man["shards"] == []
    IS.ShardedIndexWriter(tmp_path / "idx", dim=8).append(_rows(3, 8, 1))
    e, p, t = IS.load_rows(tmp_path / "idx", 3, 3)
    assert e.shape == (0, 8) and p == [] and t == []


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_writer(rank, world, port, d, n, dim, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = _rows(n, dim, 7)
        lo, hi = shard_bounds(n, rank, world)
        w = IS.ShardedIndexWriter(d, dim, order_major=rank)
        mid = (lo + hi) // 2                      # two shards per rank, written without any collective
        w.append(full[lo:mid], [f"img{i}" for i in range(lo, mid)], None)
        w.append(full[mid:hi], [f"img{i}" for i in range(mid, hi)], None)
        dist.barrier()
        if rank == 0:
            IS.write_manifest(d)
        dist.barrier()
        # every rank reloads ITS row block of the merged index: exactly the rows it would scan
        man = IS.read_manifest(d)
        e, paths, _ = IS.load_rows(d, lo, hi)
        ret[rank] = bool(man["rows"] == n and torch.equal(e, full[lo:hi]) and paths[0] == f"img{lo}")
    finally:
        dist.destroy_process_group()


def test_two_rank_data_parallel_build(tmp_path):
    port, mgr = _free_port(), mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rank_writer, args=(2, port, str(tmp_path / "idx"), 1001, 24, ret), nprocs=2, join=True)
    assert all(ret.get(r) for r in range(2)), dict(ret)
    man = json.loads((tmp_path / "idx" / "manifest.json").read_text())
    assert [tuple(s["order"]) for s in man["shards"]] == [(0, 0), (0, 1), (1, 0), (1, 1)]
    e, _, _ = IS.load_rows(tmp_path / "idx", 0, 1001)
    assert torch.equal(e, _rows(1001, 24, 7))


# ------------------------------------------------------------------------------------------
# the other two index writers of the reference (host logic; the encoder is a stub here)
# ------------------------------------------------------------------------------------------
def _stub_encode(texts):
    g = torch.Generator().manual_seed(len(texts))
    return torch.randn((len(texts), 8), generator=g) * 3.0   # not unit length on purpose


def test_build_custom_index_keeps_the_reference_column_quirk(tmp_path):
    """scripts/build_custom_index.py: header `image_path,text` over three-field rows read with index_col=0 ->
    index = image path, 'image_path' = description, 'text' = location; caption = "<description>, <location>";
    singular metadata keys; unit rows; the reference's errors."""
    from clip_lora_match_b200.scripts.build_custom_index import build_custom_index, read_custom_csv
    from clip_lora_match_b200.src.embedding.search import TextSearchIndex  # noqa: F401  (import check only)

    csv = tmp_path / "my_items.csv"
    csv.write_text("image_path,text\n"
                   "data/custom/images/a.jpg,Kaca mata pink, ditemukan di gk 1.\n"
                   "data/custom/images/b.jpg,Tas ransel hitam, ditemukan di aula gedung f.\n")
    paths, texts = read_custom_csv(csv)
    assert paths == ["data/custom/images/a.jpg", "data/custom/images/b.jpg"]
    assert texts == ["Kaca mata pink,  ditemukan di gk 1.", "Tas ransel hitam,  ditemukan di aula gedung f."]
    out = tmp_path / "idx" / "custom_items_index.pt"
    emb = build_custom_index(csv, out, _stub_encode, log=lambda m: None)
    obj = torch.load(out)
    assert set(obj) == {"embeddings", "image_path", "text"}
    assert obj["image_path"] == paths and obj["text"] == texts
    assert torch.allclose(obj["embeddings"].norm(dim=-1), torch.ones(2), atol=1e-6) and torch.equal(obj["embeddings"], emb)
    with pytest.raises(FileNotFoundError):
        read_custom_csv(tmp_path / "nope.csv")
    bad = tmp_path / "bad.csv"
    bad.write_text("path,caption\nx,y,z\n")
    with pytest.raises(ValueError):
        read_custom_csv(bad)
    empty = tmp_path / "empty.csv"
    empty.write_text("image_path,text\n")
    with pytest.raises(ValueError):
        read_custom_csv(empty)


def test_rebuild_index_orders_by_id_and_writes_plural_keys(tmp_path):
    """scripts/rebuild_index.py: items in id order, plural metadata keys (the spelling FinderService reads),
    nothing written for an empty table, records may be dicts or attribute objects."""
    from types import SimpleNamespace

    from clip_lora_match_b200.scripts.rebuild_index import read_items_jsonl, rebuild_index

    items = [SimpleNamespace(id=3, description="tas pink", image_path="c.jpg"),
             {"id": 1, "description": "kaca mata", "image_path": "a.jpg"},
             SimpleNamespace(id=2, description="sepatu", image_path="b.jpg")]
    out = tmp_path / "index.pt"
    emb = rebuild_index(items, out, _stub_encode, log=lambda m: None)
    obj = torch.load(out)
    assert set(obj) == {"embeddings", "image_paths", "texts"}
    assert obj["image_paths"] == ["a.jpg", "b.jpg", "c.jpg"] and obj["texts"] == ["kaca mata", "sepatu", "tas pink"]
    assert torch.allclose(obj["embeddings"].norm(dim=-1), torch.ones(3), atol=1e-6) and emb.shape == (3, 8)
    none_out = tmp_path / "none.pt"
    assert rebuild_index([], none_out, _stub_encode, log=lambda m: None) is None and not none_out.exists()
    dump = tmp_path / "items.jsonl"
    dump.write_text('{"id": 2, "description": "b", "image_path": "b.png"}\n\n{"id": 1, "description": "a", "image_path": "a.png"}\n')
    assert [r["id"] for r in read_items_jsonl(dump)] == [2, 1]
    with pytest.raises(FileNotFoundError):
        read_items_jsonl(tmp_path / "missing.jsonl")


def test_finder_service_report_item_appends_like_the_reference(tmp_path):
    """src/embedding/finder_service.py mirror (host logic; stub text encoder): the image is copied into
    upload_dir, only the caption "<description>, ditemukan di <location>" is embedded, the one-file index
    grows by load-cat-save with the plural keys, a directory index grows by one shard per report, both stay
    readable row for row; errors as the reference."""
    from datetime import datetime

    from clip_lora_match_b200.src.embedding.finder_service import FinderConfig, FinderService

    root = tmp_path / "svc"
    (root / "incoming").mkdir(parents=True)
    img = root / "incoming" / "tas-pink.jpg"
    img.write_bytes(b"\xff\xd8not-a-real-jpeg")
    seen = []

    def enc(text):
        seen.append(text)
        g = torch.Generator().manual_seed(len(seen))
        return torch.randn((8,), generator=g) * 2.0

    ids = iter([41, 42])
    for index_path in (root / "data" / "index" / "custom_items_index.pt", root / "data" / "index" / "sharded"):
        seen.clear()
        cfg = FinderConfig(root_dir=root, clip_config_path=root / "none.yaml", lora_dir=root / "none",
                           index_path=index_path, upload_dir=root / "data" / "reported" / "images")
        svc = FinderService(cfg, encode_fn=enc, on_item=(lambda rec: next(ids)) if index_path.suffix else None)
        r1 = svc.report_item(img, "tas pink kanken", location="lab iot", reporter="ani",
                             found_at=datetime(2025, 1, 2, 3, 4, 5))
        r2 = svc.report_item(img, "kaca mata pink")
        assert (cfg.upload_dir / "tas-pink.jpg").read_bytes() == img.read_bytes()
        assert r1["image_path"] == "data/reported/images/tas-pink.jpg" and r1["location"] == "lab iot"
        assert r1["description"] == "tas pink kanken, ditemukan di lab iot" and r2["description"] == "kaca mata pink"
        assert r1["found_at"] == "2025-01-02T03:04:05" and r1["reporter"] == "ani" and r2["found_at"] is not None
        assert (r1["id"], r2["id"]) == ((41, 42) if index_path.suffix else (1, 2))
        assert seen == ["tas pink kanken, ditemukan di lab iot", "kaca mata pink"]   # the TEXT is what is embedded
        if index_path.suffix:
            obj = torch.load(index_path)
            assert set(obj) == {"embeddings", "image_paths", "texts"}
            emb, paths, texts = obj["embeddings"], obj["image_paths"], obj["texts"]
        else:
            assert IS.read_manifest(index_path)["rows"] == 2 and len(IS.scan_shards(index_path)) == 2
            emb, paths, texts = IS.load_rows(index_path, 0, 2)
        assert emb.shape == (2, 8) and torch.allclose(emb.norm(dim=-1), torch.ones(2), atol=1e-6)
        assert paths == ["data/reported/images/tas-pink.jpg"] * 2 and texts == seen
        with pytest.raises(FileNotFoundError):
            svc.report_item(root / "incoming" / "missing.png", "x")


def test_two_writers_with_the_same_order_major_never_collide_and_manifest_appends_incrementally(tmp_path):
    """Two finder processes appending to one shard directory (same order_major) used to compute the same next
    sequence number and os.replace one report's shard with the other's; the number is now claimed with an
    exclusive create.  append_to_manifest adds one shard without re-reading the sidecars and must agree with a
    full rebuild; it falls back to the rebuild when the manifest is missing or behind."""
    d = tmp_path / "idx"
    w1 = IS.ShardedIndexWriter(d, 4)
    w2 = IS.ShardedIndexWriter(d, 4)          # constructed before w1 wrote anything: same starting number
    rows = [torch.eye(4)[i:i + 1] for i in range(4)]
    p1 = w1.append(rows[0], ["a.png"], ["a"])
    man = IS.append_to_manifest(d, w1.last_shard, w1.dim)     # no manifest yet -> full rebuild
    p2 = w2.append(rows[1], ["b.png"], ["b"])
    man = IS.append_to_manifest(d, w2.last_shard, w2.dim)
    p3 = w1.append(rows[2], ["c.png"], ["c"])
    man = IS.append_to_manifest(d, w1.last_shard, w1.dim)
    assert len({p1.name, p2.name, p3.name}) == 3
    assert man["rows"] == 3 and man == IS.write_manifest(d)
    emb, paths, texts = IS.load_rows(d, 0, 3)
    assert sorted(paths) == ["a.png", "b.png", "c.png"] and emb.shape == (3, 4)
    # a manifest that already lists a later shard (another process got ahead) -> rebuilt, not corrupted
    p4 = w2.append(rows[3], ["d.png"], ["d"])
    stale = w1.last_shard
    man = IS.append_to_manifest(d, stale, 4)
    assert man["rows"] == 4 and man == IS.write_manifest(d)
