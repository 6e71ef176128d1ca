"""Generate tests/golden/*.npz by running the REFERENCE's own code in the build container.

TEST INFRASTRUCTURE.  Needs /root/reference (read-only) — it exists only in the build
container, so the fixtures are committed and this script documents how they were made:

    python -m oracle.make_golden

What is executed from the reference, unmodified, imported from /root/reference:
  * src/embedding/search.py  TextSearchIndex.__init__ / .search_with_embedding
  * src/embedding/similarity.py  cosine_similarity / top_k_similar
  * models/lora_adapter.py  create_lora_config / attach_lora_to_clip
  * models/clip_model.py  encode_image / encode_text
with `peft` replaced by oracle/peft_stub.py (the package is not installable here) and the
transformers-5.x CLIPModel wrapped so that get_*_features return the pooled tensor as the
4.x API the reference was written against did.
"""
from __future__ import annotations

import io
import sys
import tempfile
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
GOLD = ROOT / "tests" / "golden"

sys.path.insert(0, str(ROOT))
from oracle import clip_oracle as O  # noqa: E402
from oracle import peft_stub  # noqa: E402


def import_reference():
    if not REF.exists():
        raise SystemExit("/root/reference is not present: golden vectors can only be regenerated "
                         "in the build container")
    sys.modules["peft"] = peft_stub
    sys.path.insert(0, str(REF))
    import models.clip_model as ref_cm
    import models.lora_adapter as ref_la
    import src.embedding.search as ref_search
    import src.embedding.similarity as ref_sim
    return ref_cm, ref_la, ref_search, ref_sim


class Tf4Compat(nn.Module):
    """transformers-4.x behaviour of get_image_features/get_text_features over 5.x."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def get_image_features(self, pixel_values=None, **kw):
        return O._pooled(self.model.get_image_features(pixel_values=pixel_values, **kw))

    def get_text_features(self, input_ids=None, attention_mask=None, **kw):
        return O._pooled(self.model.get_text_features(input_ids=input_ids, attention_mask=attention_mask, **kw))


def synth_png(seed: int, path: Path) -> None:
    """Seeded 224x224 RGB noise image (tests regenerate the same array)."""
    from PIL import Image

    rng = np.random.RandomState(seed)
    arr = rng.randint(0, 256, size=(224, 224, 3), dtype=np.uint8)
    Image.fromarray(arr, "RGB").save(path)


class ShimProcessor:
    """processor(images=...) -> the real offline CLIPImageProcessor(); processor(text=[...]) ->
    pre-chosen input_ids (no BPE vocabulary exists on this box)."""

    def __init__(self, ids_by_text):
        from transformers import CLIPImageProcessor

        self.ip = CLIPImageProcessor()
        self.ids_by_text = ids_by_text

    def __call__(self, text=None, images=None, return_tensors="pt", **kw):
        if images is not None:
            return self.ip(images=images, return_tensors="pt")
        ids = torch.stack([self.ids_by_text[t] for t in text])
        return {"input_ids": ids, "attention_mask": torch.ones_like(ids)}


def search_golden(ref_search, ref_sim):
    out = {}
    buf = io.StringIO()
    # (1) the one fixture the reference ships
    with redirect_stdout(buf):
        idx = ref_search.TextSearchIndex(REF / "data" / "index" / "custom_items_index.pt")
    raw = torch.load(REF / "data" / "index" / "custom_items_index.pt", map_location="cpu")
    out["fixture_embeddings"] = raw["embeddings"].float().numpy()
    g = torch.Generator().manual_seed(100)
    queries = torch.cat([raw["embeddings"].float()[:3] * 2.5, torch.randn((5, idx.dim), generator=g)], 0)
    out["fixture_queries"] = queries.numpy()
    for k in (1, 3, 5, 10):
        ids, scores = [], []
        for q in queries:
            res = idx.search_with_embedding(q, top_k=k)
            ids.append([r.index for r in res])
            scores.append([r.score for r in res])
        out[f"fixture_ids_k{k}"] = np.asarray(ids, dtype=np.int64)
        out[f"fixture_scores_k{k}"] = np.asarray(scores, dtype=np.float32)
    # (2) synthetic index through the same class (writer format of build_text_index.py:69-73)
    n, d, nq = 3000, 512, 24
    emb = O.synth_unit_rows(n, d, 4) * 1.7  # un-normalised on purpose: __init__ renormalises (:68)
    qs = torch.randn((nq, d), generator=torch.Generator().manual_seed(5))
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "idx.pt"
        torch.save({"embeddings": emb, "image_path": [f"img{i}.jpg" for i in range(n)],
                    "text": [f"t{i}" for i in range(n)]}, p)
        with redirect_stdout(buf):
            idx2 = ref_search.TextSearchIndex(p)
        ids, scores = [], []
        for q in qs:
            res = idx2.search_with_embedding(q, top_k=10)
            ids.append([r.index for r in res])
            scores.append([r.score for r in res])
    out["synth_n"], out["synth_d"], out["synth_nq"] = n, d, nq
    out["synth_ids_k10"] = np.asarray(ids, dtype=np.int64)
    out["synth_scores_k10"] = np.asarray(scores, dtype=np.float32)
    # (3) similarity.py
    tv, ti = [], []
    for q in qs[:8]:
        v, i = ref_sim.top_k_similar(q, emb, k=5)
        tv.append(v.numpy())
        ti.append(i.numpy())
    out["sim_values_k5"] = np.stack(tv)
    out["sim_indices_k5"] = np.stack(ti)
    out["sim_cosine_q0"] = ref_sim.cosine_similarity(qs[0], emb).numpy()[:64]
    np.savez_compressed(GOLD / "search_golden.npz", **out)
    print("search_golden.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})


def encoder_golden(ref_cm, ref_la):
    out = {}
    buf = io.StringIO()
    cases = [("tiny-test", 4, 4, 8, 16, ["q_proj", "v_proj"]),
             ("tiny-test", 2, 2, 8, 16, ["q_proj", "k_proj", "v_proj", "out_proj"]),
             ("openai/clip-vit-base-patch32", 2, 2, 8, 16, ["q_proj", "v_proj"])]
    for ci, (arch, n_img, n_txt, r, alpha, targets) in enumerate(cases):
        model = O.build_model(arch, seed=0)
        shim = Tf4Compat(model)
        with tempfile.TemporaryDirectory() as td:
            y = Path(td) / "lora.yaml"
            y.write_text("model:\n  target_modules:\n" + "".join(f"    - \"{t}\"\n" for t in targets) +
                         f"lora:\n  r: {r}\n  alpha: {alpha}\n  dropout: 0.1\n  bias: \"none\"\n")
            cfg = ref_la.create_lora_config(y)                      # reference code
            with redirect_stdout(buf):
                peft_model = ref_la.attach_lora_to_clip(shim, cfg)  # reference code
            trainable, total = peft_model.trainable_parameter_count()
            peft_model.eval()
            # non-zero B (PEFT's B=0 would make LoRA a no-op): same seeded weights the tests rebuild
            g = torch.Generator().manual_seed(1)
            weights = {}
            for path in O.get_lora_weights(shim).keys():
                mod = shim.get_submodule(path)
                a = torch.empty_like(mod.lora_A.weight)
                bound = 1.0 / (a.shape[1] ** 0.5)
                a.uniform_(-bound, bound, generator=g)
                b = torch.randn(mod.lora_B.weight.shape, generator=g) * 0.02
                weights[path] = (a, b)
            O.set_lora_weights(shim, weights)
            ids, _ = O.synth_captions(n_txt, seed=3)
            # reference encode_text tokenises with padding=True on ONE text => no padding: cut at first EOS
            texts, ids_by_text = [], {}
            for i in range(n_txt):
                row = ids[i]
                end = int((row == O.EOS_ID).nonzero()[0]) + 1
                texts.append(f"caption-{i}")
                ids_by_text[f"caption-{i}"] = row[:end]
            proc = ShimProcessor(ids_by_text)
            dev = torch.device("cpu")
            img_emb = []
            for i in range(n_img):
                p = Path(td) / f"img{i}.png"
                synth_png(1000 + i, p)
                img_emb.append(ref_cm.encode_image(p, peft_model, proc, dev).numpy())   # reference code
            txt_emb = [ref_cm.encode_text(t, peft_model, proc, dev).numpy() for t in texts]  # reference code
        pre = f"case{ci}_"
        out[pre + "arch"] = arch
        out[pre + "targets"] = ",".join(targets)
        out[pre + "r"], out[pre + "alpha"] = r, alpha
        out[pre + "n_wrapped"] = len(weights)
        out[pre + "trainable"], out[pre + "total"] = trainable, total
        out[pre + "image_seeds"] = np.arange(1000, 1000 + n_img)
        out[pre + "input_ids"] = ids.numpy()
        out[pre + "image_emb"] = np.stack(img_emb)
        out[pre + "text_emb"] = np.stack(txt_emb)
        print(f"case {ci}: {arch} wrapped={len(weights)} trainable={trainable}")
    out["n_cases"] = len(cases)
    np.savez_compressed(GOLD / "encoder_golden.npz", **out)


def fusion_golden():
    """Run the reference's OWN SeekerService._build_query_embedding (src/embedding/seeker_service.py:84-157)
    on given text / image embeddings: `ultralytics` (YOLO) is absent, so yolo_cropper is stubbed; the
    module's encode_text / encode_image are replaced by table look-ups so that only the fusion
    arithmetic under test runs (it is the unmodified reference method)."""
    import types

    stub = types.ModuleType("src.preprocessing.yolo_cropper")
    stub.YoloCropper = object
    sys.modules["src.preprocessing.yolo_cropper"] = stub
    import src.embedding.seeker_service as ref_ss

    d, n = 64, 12
    txt = O.synth_unit_rows(n, d, 11)
    img = O.synth_unit_rows(n, d, 12)
    table_t = {f"q{i}": txt[i] for i in range(n)}
    table_i = {Path(f"/img{i}.png"): img[i] for i in range(n)}
    ref_ss.encode_text = lambda text, model, processor, device: table_t[text].clone()
    ref_ss.encode_image = lambda path, model, processor, device: table_i[Path(path)].clone()
    svc = ref_ss.SeekerService.__new__(ref_ss.SeekerService)
    svc.model = svc.processor = svc.device = None
    svc.yolo_cropper = None
    svc.yolo_crop_dir = None
    weights = [(0.5, 0.5), (0.7, 0.3), (1.0, 0.25)]
    both = np.stack([np.stack([svc._build_query_embedding(f"q{i}", Path(f"/img{i}.png"), wt, wi).numpy()
                               for i in range(n)]) for wt, wi in weights])
    only_t = np.stack([svc._build_query_embedding(f"q{i}", None).numpy() for i in range(n)])
    only_i = np.stack([svc._build_query_embedding("  ", Path(f"/img{i}.png")).numpy() for i in range(n)])
    try:
        svc._build_query_embedding(None, None)
        raised = ""
    except ValueError as e:
        raised = str(e)
    np.savez_compressed(GOLD / "fusion_golden.npz", text=txt.numpy(), image=img.numpy(),
                        weights=np.array(weights, dtype=np.float32), both=both, only_text=only_t,
                        only_image=only_i, error=np.array(raised))
    print(f"fusion golden: {both.shape} fused rows, error message {raised!r}")


def train_golden():
    """Run the reference's OWN compute_clip_contrastive_loss (scripts/train_lora.py:83-108) with autograd on
    seeded features, and its OWN lr_lambda closure (train_lora.py:148-151; the nested function is compiled from
    the reference's source, unmodified, with total_steps / warmup_steps bound as in train()).  `datasets.dataset`
    (the CSV / PIL loader the script imports at module level) is stubbed: it needs the repo's working directory
    layout and is not part of the arithmetic under test."""
    import ast
    import types

    stub = types.ModuleType("datasets.dataset")
    stub.ClipPairDataset = object
    pkg = types.ModuleType("datasets")
    pkg.dataset = stub
    sys.modules["datasets"], sys.modules["datasets.dataset"] = pkg, stub
    import scripts.train_lora as ref_tl

    out = {}
    cases = [(8, 512, 0.07), (32, 512, 0.07), (5, 64, 0.07), (16, 768, 0.2), (64, 128, 0.05)]
    for ci, (n, d, temp) in enumerate(cases):
        g = torch.Generator().manual_seed(100 + ci)
        fi = (torch.randn((n, d), generator=g) * 3.0).requires_grad_(True)
        # correlated text features so that the loss is not just log(n)
        ft = (0.1 * fi.detach() + torch.randn((n, d), generator=g) * 2.0).requires_grad_(True)
        loss = ref_tl.compute_clip_contrastive_loss(fi, ft, temp)  # reference code
        loss.backward()
        out[f"c{ci}_shape"] = np.array([n, d], dtype=np.int64)
        out[f"c{ci}_temp"] = np.array(temp, dtype=np.float64)
        out[f"c{ci}_fi"], out[f"c{ci}_ft"] = fi.detach().numpy(), ft.detach().numpy()
        out[f"c{ci}_loss"] = np.array(loss.item(), dtype=np.float64)
        out[f"c{ci}_dfi"], out[f"c{ci}_dft"] = fi.grad.numpy(), ft.grad.numpy()
    out["n_cases"] = np.array(len(cases))
    # the schedule closure, from the reference's source
    src = (REF / "scripts" / "train_lora.py").read_text()
    fn = [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "lr_lambda"][0]
    sched = []
    for total, ratio in [(100, 0.1), (7, 0.1), (50, 0.0), (20, 0.5)]:
        env = {"total_steps": total, "warmup_steps": int(total * ratio)}  # train_lora.py:144-146
        exec(compile(ast.Module(body=[fn], type_ignores=[]), "train_lora.py", "exec"), env)
        sched.append([total, env["warmup_steps"]] + [env["lr_lambda"](s) for s in range(total + 2)] + [0.0] * (100 - total))
    out["sched"] = np.array(sched, dtype=np.float64)
    np.savez_compressed(GOLD / "train_golden.npz", **out)
    print(f"train golden: {len(cases)} loss cases, losses {[float(out[f'c{i}_loss']) for i in range(len(cases))]}")


def main():
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    ref_cm, ref_la, ref_search, ref_sim = import_reference()
    if "--train-only" in sys.argv:
        train_golden()
        return
    if "--fusion-only" not in sys.argv:
        search_golden(ref_search, ref_sim)
        encoder_golden(ref_cm, ref_la)
    fusion_golden()
    train_golden()


if __name__ == "__main__":
    main()
