"""Every reference-facing function of SURVEY.md §8(a), driven the way the reference's callers drive it, on the
GPU, each compared with the CPU oracle (never with the GPU path itself):

  load_clip_model (YAML -> local checkpoint -> LoRA dir)    reference models/clip_model.py:37-82
  encode_image / encode_text                                reference models/clip_model.py:89-150
  create_lora_config + attach_lora_to_clip                  reference models/lora_adapter.py:21-56
  embed_image / embed_images_batch / embed_text             reference src/embedding/embed_image.py:22-98, embed_text.py:11-60
  TextSearchIndex.search_by_text / search_by_image          reference src/embedding/search.py:117-151
  scripts/build_text_index.main -> TextSearchIndex          reference scripts/build_text_index.py:52-75

The checkpoint is the oracle's tiny random-init CLIPModel written with save_pretrained, so load_clip_model runs its
real from_pretrained path (no network needed) and the oracle holds exactly the same weights.
"""
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle as O

pytestmark = pytest.mark.gpu

COS_MIN = 0.999
TARGETS = ("q_proj", "k_proj", "v_proj", "out_proj")  # the shipped lora_config.yaml's targets


def _write_yaml(path, name, lora_dir=None, extra=""):
    lines = ["model:", f'  name: "{name}"', '  device: "cuda"', '  dtype: "bfloat16"', extra, "paths:"]
    lines.append(f'  lora_weights_dir: "{lora_dir}"' if lora_dir else "  checkpoints_dir: \"unused\"")
    path.write_text("\n".join(l for l in lines if l) + "\n")
    return path


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.float().cpu(), b.float().cpu(), dim=-1)


def _images(n, seed=11):
    from PIL import Image

    rs = np.random.RandomState(seed)
    return [Image.fromarray(rs.randint(0, 256, size=(200 + 13 * i, 180 + 7 * i, 3), dtype=np.uint8), "RGB")
            for i in range(n)]


CAPTIONS = ["tas pink kanken, ditemukan di lab iot", "dompet kulit coklat", "botol minum biru tua dengan stiker",
            "kunci motor honda gantungan merah", "payung lipat hitam", "a", "kacamata frame bulat emas di kantin",
            "jaket denim ukuran l tertinggal di perpustakaan lantai dua dekat jendela besar"]


@pytest.fixture(scope="module")
def tiny(tmp_path_factory, cuda_device):
    """Tiny checkpoint on disk + LoRA adapter dir (PEFT layout) + YAML, and the oracle holding the same weights."""
    from clip_lora_match_b200.models.lora_adapter import LoraAdapter, LoraConfig, save_lora_adapter

    root = tmp_path_factory.mktemp("surface")
    oracle = O.build_model("tiny-test", seed=0)
    ckpt = root / "tiny-clip"
    oracle.save_pretrained(str(ckpt))
    weights = O.synthetic_lora(oracle, 8, 16, TARGETS, seed=1)  # injects the same adapter into the oracle
    lora_dir = save_lora_adapter(LoraAdapter(LoraConfig(r=8, lora_alpha=16, target_modules=list(TARGETS)), weights,
                                             base_model_name_or_path=str(ckpt)), root / "lora" / "epoch_1")
    yaml_path = _write_yaml(root / "clip_config.yaml", ckpt, lora_dir)
    return {"root": root, "oracle": oracle, "ckpt": ckpt, "lora_dir": lora_dir, "yaml": yaml_path}


@pytest.fixture(scope="module")
def loaded(tiny):
    from clip_lora_match_b200.models import clip_model as CM

    model, processor, device = CM.load_clip_model(tiny["yaml"], use_lora=True, lora_weights_path=tiny["lora_dir"])
    assert device.type == "cuda" and model.lora is not None and len(model.lora.weights) == 4 * (2 + 2)
    return model, processor, device


def _oracle_text(oracle, processor, texts):
    enc = processor(text=list(texts), return_tensors="pt", padding=True, truncation=True)
    return O.encode_texts(oracle, enc["input_ids"], enc["attention_mask"])


def _oracle_images(oracle, processor, images, normalize=True):
    pv = processor(images=images, return_tensors="pt")["pixel_values"]
    return O.encode_images(oracle, pv, normalize=normalize)


def test_load_clip_model_yaml_checkpoint_lora_then_encode(tiny, loaded, tmp_path):
    """YAML -> from_pretrained(local dir) -> PEFT-layout adapter dir -> encode_image / encode_text (one item, CPU
    fp32 (d,) out) against the oracle with the same base weights and adapter."""
    from clip_lora_match_b200.models import clip_model as CM

    model, processor, device = loaded
    img = _images(1)[0]
    p = tmp_path / "query.png"
    img.save(p)
    e = CM.encode_image(p, model, processor, device)
    assert e.shape == (64,) and e.dtype == torch.float32 and e.device.type == "cpu"
    assert abs(float(e.norm()) - 1.0) < 1e-5
    from PIL import Image

    ref = _oracle_images(tiny["oracle"], processor, [Image.open(p).convert("RGB")])[0]
    assert _cos(e, ref) >= COS_MIN
    t = CM.encode_text(CAPTIONS[0], model, processor, device)
    assert t.shape == (64,) and _cos(t, _oracle_text(tiny["oracle"], processor, [CAPTIONS[0]])[0]) >= COS_MIN
    # use_lora=True with a missing directory: the reference prints and continues WITHOUT LoRA (:70-75)
    base, _, _ = CM.load_clip_model(tiny["yaml"], use_lora=True, lora_weights_path=tiny["root"] / "nope")
    assert base.lora is None
    # lora_weights_path=None falls back to paths.lora_weights_dir of the YAML (:65-68)
    m2, _, _ = CM.load_clip_model(tiny["yaml"], use_lora=True)
    assert m2.lora is not None
    assert torch.allclose(CM.encode_text(CAPTIONS[0], m2, processor, device), t, atol=1e-6)
    # and the adapter matters: base-model embedding differs
    assert (CM.encode_text(CAPTIONS[0], base, processor, device) - t).abs().max() > 1e-4


def test_load_clip_model_raises_on_a_broken_checkpoint(tiny, tmp_path):
    """A checkpoint that exists but cannot be loaded raises (reference: from_pretrained's error propagates);
    random init is only for 'nothing on disk' or an explicit request."""
    from clip_lora_match_b200.models import clip_model as CM

    bad = tmp_path / "broken-clip"
    bad.mkdir()
    (bad / "config.json").write_text((tiny["ckpt"] / "config.json").read_text())
    (bad / "model.safetensors").write_bytes(b"not a safetensors file")
    with pytest.raises(Exception):
        CM.load_clip_model(_write_yaml(tmp_path / "bad.yaml", bad))
    with pytest.raises(FileNotFoundError):
        CM.load_clip_model(tmp_path / "missing.yaml")
    # unknown architecture name with nothing on disk: an error, not a silently different model
    with pytest.raises(ValueError):
        CM.load_clip_model(_write_yaml(tmp_path / "unk.yaml", "someone/unknown-clip"))


def test_legacy_eos_token_id_2_pools_at_the_end_of_text_token(cuda_device, tmp_path):
    """openai/clip-vit-* checkpoints carry text_config.eos_token_id = 2; transformers then pools at
    input_ids.argmax(-1), i.e. at the 49407 token (TF:575-590).  Loading such a checkpoint must pool there too
    (not at position 0) and pad with 49407."""
    from transformers import CLIPModel

    from clip_lora_match_b200.models import clip_model as CM

    cfg = O.hf_config("tiny-test")
    cfg.text_config.eos_token_id = 2
    cfg._attn_implementation = "eager"
    torch.manual_seed(5)
    hf = CLIPModel(cfg).eval().float()
    ckpt = tmp_path / "legacy-clip"
    hf.save_pretrained(str(ckpt))
    model, processor, device = CM.load_clip_model(_write_yaml(tmp_path / "legacy.yaml", ckpt))
    assert model.arch.eos_id == 49407
    ids, mask = O.synth_captions(12, seed=9)
    ref = O.encode_texts(hf, ids, mask)
    got = model.encode_texts(ids).cpu()
    assert _cos(got, ref).min() >= COS_MIN
    # captions differ, so must their embeddings (pooling at BOS would make all rows identical)
    assert (got[0] - got[1]).abs().max() > 1e-3
    # the single-caption reference call: unpadded ids of length L < 77
    t = CM.encode_text(CAPTIONS[1], model, processor, device)
    enc = processor(text=[CAPTIONS[1]], return_tensors="pt", padding=True, truncation=True)
    assert _cos(t, O.encode_texts(hf, enc["input_ids"], enc["attention_mask"])[0]) >= COS_MIN


def test_attach_lora_to_clip_and_create_lora_config(tiny, cuda_device, tmp_path, capsys):
    """attach_lora_to_clip(model, create_lora_config(yaml)): PEFT's get_peft_model semantics -- every matching
    Linear of BOTH towers wrapped, A random, B = 0 (output unchanged), trainable-parameter line printed -- then
    the same adapter with non-zero B against the oracle."""
    from clip_lora_match_b200.models import clip_model as CM
    from clip_lora_match_b200.models.lora_adapter import attach_lora_to_clip, create_lora_config

    (tmp_path / "lora.yaml").write_text(
        "model:\n  target_modules: [q_proj, k_proj, v_proj, out_proj]\nlora:\n  r: 4\n  alpha: 8\n")
    cfg = create_lora_config(tmp_path / "lora.yaml")
    assert (cfg.r, cfg.lora_alpha, cfg.lora_dropout, cfg.bias) == (4, 8, 0.1, "none")
    model, processor, device = CM.load_clip_model(_write_yaml(tmp_path / "c.yaml", tiny["ckpt"]))
    pv = O.synth_images(3, seed=2)
    ids, mask = O.synth_captions(5, seed=3)
    base_i, base_t = model.encode_images(pv).cpu(), model.encode_texts(ids).cpu()
    out = attach_lora_to_clip(model, cfg)
    assert out is model and model.lora is not None
    printed = capsys.readouterr().out
    n_train = 16 * (4 * 128 + 128 * 4)  # 4 layers (2 + 2) x 4 targets x (A [4,128] + B [128,4])
    assert f"trainable params: {n_train:,d}" in printed
    assert sorted(model.lora.weights) == sorted(O.target_paths(O.build_model("tiny-test", seed=0), TARGETS))
    assert torch.equal(model.encode_images(pv).cpu(), base_i) or \
        (model.encode_images(pv).cpu() - base_i).abs().max() < 1e-6  # B = 0: a no-op
    # train-like update of B, then parity against the oracle carrying the same adapter
    g = torch.Generator().manual_seed(3)
    adapter = model.lora
    adapter.weights = {p: (a, torch.randn(b.shape, generator=g) * 0.02) for p, (a, b) in adapter.weights.items()}
    model.set_lora(adapter)
    oracle = O.build_model("tiny-test", seed=0)
    O.inject_lora(oracle, cfg.r, cfg.lora_alpha, cfg.target_modules)
    O.set_lora_weights(oracle, adapter.weights)
    got_i, got_t = model.encode_images(pv).cpu(), model.encode_texts(ids).cpu()
    assert _cos(got_i, O.encode_images(oracle, pv)).min() >= COS_MIN
    assert _cos(got_t, O.encode_texts(oracle, ids, mask)).min() >= COS_MIN
    assert (got_i - base_i).abs().max() > 1e-4 and (got_t - base_t).abs().max() > 1e-4


def test_embed_image_embed_images_batch_embed_text(tiny, loaded, tmp_path):
    """The batched embedding surface (unused by the reference's own scripts, kept by contract)."""
    from clip_lora_match_b200.src.embedding.embed_image import embed_image, embed_images_batch
    from clip_lora_match_b200.src.embedding.embed_text import embed_text

    model, processor, device = loaded
    oracle = tiny["oracle"]
    imgs = _images(20)
    ref = _oracle_images(oracle, processor, imgs)
    one = embed_image(model, processor, imgs[3], device)
    assert one.shape == (64,) and one.device.type == "cpu" and _cos(one, ref[3]) >= COS_MIN
    path = tmp_path / "im.png"
    imgs[4].save(path)
    assert _cos(embed_image(model, processor, str(path), device), ref[4]) >= COS_MIN
    raw = embed_image(model, processor, imgs[3], device, normalize=False)
    ref_raw = _oracle_images(oracle, processor, [imgs[3]], normalize=False)[0]
    assert (raw - ref_raw).norm() / ref_raw.norm() <= 0.03 and abs(float(raw.norm()) - 1.0) > 1e-3
    batch = embed_images_batch(model, processor, imgs, device, batch_size=16)  # 16 + 4
    assert batch.shape == (20, 64) and batch.device.type == "cpu" and _cos(batch, ref).min() >= COS_MIN
    empty = embed_images_batch(model, processor, [], device)
    assert empty.numel() == 0 and empty.shape == torch.empty(0).shape  # reference :95-96
    with pytest.raises(FileNotFoundError):
        embed_image(model, processor, tmp_path / "nope.jpg", device)
    # text: str -> (d,), list -> (N, d), padded to the longest caption of the batch (:35-41)
    ref_t = _oracle_text(oracle, processor, CAPTIONS)
    t1 = embed_text(model, processor, CAPTIONS[2], device)
    assert t1.shape == (64,) and _cos(t1, ref_t[2]) >= COS_MIN
    tn = embed_text(model, processor, CAPTIONS, device)
    assert tn.shape == (len(CAPTIONS), 64) and tn.device.type == "cpu" and _cos(tn, ref_t).min() >= COS_MIN
    traw = embed_text(model, processor, CAPTIONS[:2], device, normalize=False)
    assert (traw.norm(dim=-1) - 1.0).abs().min() > 1e-3


def test_search_by_text_and_search_by_image(tiny, loaded, tmp_path):
    """index.search_by_text(q, model, processor, device, top_k) / search_by_image: encode + search in one call."""
    from clip_lora_match_b200.src.embedding.search import SearchResult, TextSearchIndex

    model, processor, device = loaded
    oracle = tiny["oracle"]
    imgs = _images(12, seed=21)
    emb = _oracle_images(oracle, processor, imgs)
    paths = [f"data/img_{i}.jpg" for i in range(12)]
    texts = [f"item {i}" for i in range(12)]
    idx_path = tmp_path / "idx.pt"
    torch.save({"embeddings": emb, "image_paths": paths, "texts": texts}, idx_path)
    index = TextSearchIndex(idx_path, device=device, verbose=False)
    qp = tmp_path / "q.png"
    imgs[5].save(qp)
    res = index.search_by_image(qp, model, processor, device, top_k=3)
    assert all(isinstance(r, SearchResult) for r in res) and len(res) == 3
    from PIL import Image

    q_ref = _oracle_images(oracle, processor, [Image.open(qp).convert("RGB")])
    ref_s, ref_i = O.search_topk(emb, q_ref, 3)
    assert res[0].index == 5 and res[0].image_path == paths[5] and res[0].text == texts[5]
    sims = O.normalize_rows(q_ref) @ O.normalize_rows(emb).T
    got_i = torch.tensor([[r.index for r in res]])
    assert O.ids_match_with_ties(ref_s, ref_i, got_i, sims)
    assert np.allclose([r.score for r in res], torch.gather(sims, 1, got_i)[0].numpy(), atol=5e-3)
    rt = index.search_by_text(CAPTIONS[0], model, processor, device, top_k=4)
    qt = _oracle_text(oracle, processor, [CAPTIONS[0]])
    ref_s, ref_i = O.search_topk(emb, qt, 4)
    sims = O.normalize_rows(qt) @ O.normalize_rows(emb).T
    got_i = torch.tensor([[r.index for r in rt]])
    # the query embedding differs from the oracle's in bf16-level digits: compare by oracle score of the ids
    assert (torch.gather(sims, 1, got_i) - ref_s).abs().max() <= 5e-3
    assert len(rt) == 4 and rt == sorted(rt, key=lambda r: -r.score)
    assert index.search_by_text(CAPTIONS[0], model, processor, device, top_k=100)[0].index == rt[0].index  # k > N


def test_build_text_index_main_then_reload(tiny, tmp_path):
    """scripts/build_text_index.main on a 20-row CSV (the reference's CLI), reloaded through TextSearchIndex;
    rows are the oracle's embeddings of the same captions, unit norm, keys as the reference writes them."""
    import pandas as pd

    from clip_lora_match_b200.models import clip_model as CM
    from clip_lora_match_b200.scripts import build_text_index as B
    from clip_lora_match_b200.src.embedding.search import TextSearchIndex

    texts = [f"{CAPTIONS[i % len(CAPTIONS)]} nomor {i}" for i in range(20)]
    csv = tmp_path / "train.csv"
    pd.DataFrame({"image_path": [f"img/{i}.jpg" for i in range(20)], "text": texts}).to_csv(csv, index=False)
    out = tmp_path / "index" / "text_index.pt"
    B.main(["--data-csv", str(csv), "--lora-dir", str(tiny["lora_dir"]), "--clip-config", str(tiny["yaml"]),
            "--index-path", str(out), "--batch-size", "8"])
    obj = torch.load(out, map_location="cpu")
    assert set(obj) == {"embeddings", "image_path", "text"} and obj["text"] == texts
    emb = obj["embeddings"]
    assert emb.shape == (20, 64) and emb.dtype == torch.float32
    assert torch.allclose(emb.norm(dim=-1), torch.ones(20), atol=1e-5)
    _, processor, _ = CM.load_clip_model(tiny["yaml"])
    ref = _oracle_text(tiny["oracle"], processor, texts)
    assert _cos(emb, ref).min() >= COS_MIN
    index = TextSearchIndex(out, device="cuda", verbose=False)
    hit = index.search_with_embedding(ref[7], top_k=1)[0]
    assert hit.index == 7 and hit.text == texts[7] and hit.image_path == "img/7.jpg"
    with pytest.raises(FileNotFoundError):
        B.build_text_index(tmp_path / "none.csv", out, tiny["yaml"])
