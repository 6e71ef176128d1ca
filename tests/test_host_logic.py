"""CPU tests of the host-side mirror: C-ABI library loads and exports every symbol that
include/clm_b200.h declares, LoRA config / matching / PEFT on-disk layout, architecture
tables, tokenizer stand-in, index-file handling.  No compute call is made (no GPU here).
"""
import json
import os
import re

import pytest
import torch

from clip_lora_match_b200 import _lib
from clip_lora_match_b200.models import clip_model as CM
from clip_lora_match_b200.models import lora_adapter as LA

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "clm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/clm_b200.h but not exported"
    # and the ctypes table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names
    assert lib.clm_version() >= 100


def test_invalid_arguments_return_error_codes_not_crashes():
    lib = _lib.load()
    rc = lib.clm_layernorm(None, None, None, None, 4, 768, 1e-5, None)
    assert rc != 0 and b"clm_layernorm" in lib.clm_last_error()
    rc = lib.clm_search_topk(None, None, 1, 1, 512, 10, 10, 1, None, None, None, 0.0, None, None, None)
    assert rc != 0
    with pytest.raises(_lib.ClmError):
        _lib.check(rc, "clm_search_topk")
    assert lib.clm_search_num_splits(4096, 1250000) >= 1


def test_cpu_tensors_are_rejected_loudly():
    from clip_lora_match_b200 import kernels as K

    with pytest.raises(ValueError, match="CUDA"):
        K.layernorm(torch.zeros(2, 512), torch.ones(512), torch.zeros(512))
    with pytest.raises(ValueError):
        CM.B200ClipModel(CM.ARCHS["openai/clip-vit-base-patch32"], {}, device="cpu")


def test_struct_layout_matches_header():
    import ctypes as C

    assert C.sizeof(_lib.TowerConfig) == 16 * 4
    assert C.sizeof(_lib.LayerWeights) == 20 * 8
    assert C.sizeof(_lib.TowerWeights) == 9 * 8


# ---- LoRA -----------------------------------------------------------------------------------
def test_create_lora_config_defaults_and_yaml(tmp_path):
    y = tmp_path / "lora.yaml"
    y.write_text("lora: {}\nmodel: {}\n")
    c = LA.create_lora_config(y)
    assert (c.r, c.lora_alpha, c.lora_dropout, c.bias) == (8, 16, 0.1, "none")
    assert c.target_modules == ["q_proj", "v_proj"] and c.task_type == "FEATURE_EXTRACTION"
    shipped = LA.create_lora_config(os.path.join(ROOT, "clip_lora_match_b200", "config", "lora_config.yaml"))
    assert shipped.target_modules == ["q_proj", "k_proj", "v_proj", "out_proj"]
    assert shipped.scaling == 2.0
    with pytest.raises(FileNotFoundError):
        LA.create_lora_config(tmp_path / "missing.yaml")


def test_target_matching_hits_both_towers():
    names = LA.linear_module_paths(vision_layers=12, text_layers=12)
    qv = LA.match_target_modules(names, ["q_proj", "v_proj"])
    assert len(qv) == 48  # 24 attention blocks x 2 (golden: case2 n_wrapped)
    assert any(n.startswith("text_model.") for n in qv) and any(n.startswith("vision_model.") for n in qv)
    assert len(LA.match_target_modules(names, ["q_proj", "k_proj", "v_proj", "out_proj"])) == 96
    assert LA.match_target_modules(names, ["proj"]) == []  # suffix must follow a dot
    assert LA.match_target_modules(names, ["visual_projection"]) == ["visual_projection"]


def _dims(width_v=768, width_t=512, layers=2):
    d = {}
    for tower, w in (("text_model", width_t), ("vision_model", width_v)):
        for i in range(layers):
            for m in ("q_proj", "k_proj", "v_proj", "out_proj"):
                d[f"{tower}.encoder.layers.{i}.self_attn.{m}"] = (w, w)
            d[f"{tower}.encoder.layers.{i}.mlp.fc1"] = (4 * w, w)
    return d


def test_adapter_init_export_roundtrip(tmp_path):
    cfg = LA.LoraConfig(r=8, lora_alpha=16, target_modules=["q_proj", "v_proj"])
    ad = LA.init_lora_adapter(_dims(), cfg, seed=1, init_b_std=0.02, base_model_name="openai/clip-vit-base-patch32")
    assert len(ad.weights) == 8
    a, b = ad.weights["vision_model.encoder.layers.0.self_attn.q_proj"]
    assert a.shape == (8, 768) and b.shape == (768, 8)
    assert a.abs().max() <= 1 / 768 ** 0.5 + 1e-7
    zero_b = LA.init_lora_adapter(_dims(), cfg, seed=1)
    assert all(bb.abs().max() == 0 for _, bb in zero_b.weights.values())  # PEFT default
    out = LA.save_lora_adapter(ad, tmp_path / "epoch_1")
    assert (out / "adapter_config.json").exists() and (out / "adapter_model.safetensors").exists()
    raw = json.loads((out / "adapter_config.json").read_text())
    assert raw["peft_type"] == "LORA" and raw["r"] == 8 and raw["lora_alpha"] == 16
    from safetensors.torch import load_file

    keys = set(load_file(str(out / "adapter_model.safetensors")).keys())
    assert "base_model.model.text_model.encoder.layers.1.self_attn.v_proj.lora_B.weight" in keys
    back = LA.load_lora_adapter(out)
    assert back.config.r == 8 and back.scaling == 2.0 and set(back.weights) == set(ad.weights)
    for k in ad.weights:
        assert torch.equal(back.weights[k][0], ad.weights[k][0])
        assert torch.equal(back.weights[k][1], ad.weights[k][1])


def test_adapter_layout_interoperates_with_peft_semantics(tmp_path):
    """An adapter saved by the oracle's peft stand-in (the layout the reference's
    train_lora.py:243-247 writes) loads here, and vice versa."""
    from oracle import clip_oracle as O
    from oracle import peft_stub

    model = O.build_model("tiny-test", seed=0)
    pm = peft_stub.get_peft_model(model, peft_stub.LoraConfig(r=8, lora_alpha=16,
                                                              target_modules=["q_proj", "v_proj"]))
    pm.save_pretrained(tmp_path / "a")
    ad = LA.load_lora_adapter(tmp_path / "a")
    assert len(ad.weights) == 8 and ad.config.target_modules == ["q_proj", "v_proj"]
    LA.save_lora_adapter(ad, tmp_path / "b")
    model2 = O.build_model("tiny-test", seed=0)
    pm2 = peft_stub.PeftModel.from_pretrained(model2, str(tmp_path / "b"))
    w1, w2 = O.get_lora_weights(model), O.get_lora_weights(model2)
    assert set(w1) == set(w2)
    for k in w1:
        assert torch.equal(w1[k][0], w2[k][0])
    assert pm2.peft_config.r == 8


def test_unsupported_lora_targets_fail_loudly():
    # every Linear of the encoder layers is a valid target (fc1 / fc2 included) ...
    mlp = LA.init_lora_adapter(_dims(), LA.LoraConfig(r=4, target_modules=["fc1", "fc2"]))
    assert all(p.endswith(("mlp.fc1", "mlp.fc2")) for p in mlp.weights) and len(mlp.weights) > 0
    a, b = next(v for k, v in mlp.weights.items() if k.endswith("fc1"))
    assert a.shape[0] == 4 and b.shape[1] == 4 and b.shape[0] == 4 * a.shape[1]
    # ... the two projection heads are not
    d = dict(_dims())
    d["visual_projection"] = (64, 128)
    with pytest.raises(NotImplementedError):
        LA.init_lora_adapter(d, LA.LoraConfig(target_modules=["visual_projection"]))
    with pytest.raises(FileNotFoundError):
        LA.load_lora_adapter("/nonexistent/adapter")


# ---- architectures / processor ---------------------------------------------------------------
def test_arch_tables_agree_with_transformers_configs():
    from oracle import clip_oracle as O

    for name in ("openai/clip-vit-base-patch32", "openai/clip-vit-base-patch16", "openai/clip-vit-large-patch14"):
        arch = CM.arch_from_name(name)
        back = CM.arch_from_hf_config(CM.hf_config_for(arch), name)
        assert back == arch
        vw, vl, vh, vm, p, tw, tl, th, tm, proj = O.ARCH_TABLE[name]
        assert (arch.vision.width, arch.vision.layers, arch.vision.heads, arch.vision.mlp, arch.patch) == (vw, vl, vh, vm, p)
        assert (arch.text.width, arch.text.layers, arch.text.heads, arch.text.mlp, arch.proj_dim) == (tw, tl, th, tm, proj)
    assert CM.ARCHS["openai/clip-vit-base-patch32"].vision_tokens == 50
    assert CM.ARCHS["openai/clip-vit-base-patch16"].vision_tokens == 197
    assert CM.ARCHS["openai/clip-vit-large-patch14"].vision_tokens == 257
    with pytest.raises(ValueError):
        CM.arch_from_name("openai/clip-vit-huge")


def test_fallback_tokenizer_framing():
    tok = CM.FallbackTokenizer()
    enc = tok(["tas pink kanken, ditemukan di lab iot", "kaca mata"], padding=True, truncation=True)
    ids, mask = enc["input_ids"], enc["attention_mask"]
    assert ids.shape == mask.shape and ids.shape[0] == 2
    assert (ids[:, 0] == CM.BOS_ID).all()
    for row, m in zip(ids, mask):
        n = int(m.sum())
        assert row[n - 1] == CM.EOS_ID and (row[n:] == CM.EOS_ID).all()
        assert (row[1:n - 1] < CM.BOS_ID).all()
    long = tok(["x " * 200], padding=True, truncation=True)["input_ids"]
    assert long.shape[1] == 77 and long[0, -1] == CM.EOS_ID
    assert tok(["a"], padding="max_length")["input_ids"].shape[1] == 77
    # deterministic
    assert torch.equal(tok(["same text"])["input_ids"], tok(["same text"])["input_ids"])


def test_config_errors_match_reference_behaviour(tmp_path):
    with pytest.raises(FileNotFoundError):
        CM.load_clip_model(tmp_path / "missing.yaml")
    y = tmp_path / "clip.yaml"
    y.write_text("model:\n  name: openai/clip-vit-base-patch32\n  device: cpu\n")
    with pytest.raises(ValueError, match="CUDA"):
        CM.load_clip_model(y)


def test_shard_bounds_partition_rows():
    from clip_lora_match_b200.src.embedding.search import shard_bounds

    for n, w in ((10_000_000, 8), (6, 4), (7, 8), (1, 1)):
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(10_000_000, 3, 8) == (3_750_000, 5_000_000)


def test_text_length_buckets_cover_every_caption_and_merge_small_groups():
    """encode_texts' bucket policy (host logic, no GPU): every caption's pass is at least as long as the caption,
    a multiple of 16 positions (or the full width), and groups too small to leave the launch-bound regime are
    merged upwards."""
    import torch

    from clip_lora_match_b200.models.clip_model import B200ClipModel as M

    g = torch.Generator().manual_seed(0)
    lens = torch.randint(3, 78, (1024,), generator=g)
    tops = M._bucket_tops(M, lens, 77)
    assert bool((tops >= lens).all()) and set(tops.tolist()) <= {16, 32, 48, 64, 77}
    for t in sorted(set(tops.tolist()))[:-1]:  # every pass but the last is out of the launch-bound regime
        assert int((tops == t).sum()) * t >= M.BUCKET_MIN_TOKENS
    assert set(M._bucket_tops(M, torch.full((4096,), 12), 77).tolist()) == {16}          # uniformly short: one short pass
    assert set(M._bucket_tops(M, torch.randint(3, 78, (100,), generator=g), 77).tolist()) == {77}  # too few rows to split
    ragged = torch.tensor([5] * 3000 + [77] * 2)   # a tiny tail group is still served (at the full width)
    tr = M._bucket_tops(M, ragged, 77)
    assert bool((tr >= ragged).all()) and set(tr[:3000].tolist()) == {16} and set(tr[3000:].tolist()) == {77}


@pytest.mark.parametrize("name", ["build_text_index", "build_custom_index", "rebuild_index", "build_image_index",
                                  "demo_search_text", "demo_search_image", "demo_seeker", "demo_finder_report",
                                  "train_lora"])
def test_every_script_mirror_imports_and_parses_its_arguments(name):
    """The reference's scripts have hard-coded paths; the mirrors take them as arguments.  Importing a script
    must not need a GPU, and --help must describe it."""
    import importlib

    mod = importlib.import_module(f"clip_lora_match_b200.scripts.{name}")
    with pytest.raises(SystemExit) as e:
        mod.main(["--help"])
    assert e.value.code == 0


def test_built_library_contains_tcgen05_and_tma_sass():
    """The hot kernels are tcgen05/TMEM + TMA code, not mma.sync recompiled: the built sm_100a library must
    carry the SASS mnemonics of /opt/skills/guides/B200_PROFILING.md (tcgen05.mma -> UTC*MMA incl. the
    cta_group::2 form, tcgen05.ld/st -> LDTM/STTM, TMA loads / stores / reduce-add -> UTMALDG / UTMASTG /
    UTMAREDG) and no legacy tensor-core instruction."""
    import shutil
    import subprocess

    from clip_lora_match_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    _lib.load()
    sass = subprocess.run([cuobjdump, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG.2D", "UTMALDG.2D.2CTA", "UTMASTG.2D",
                     "UTMAREDG.2D.ADD", "UTCBAR"):
        assert mnemonic in sass, f"{mnemonic} missing from the built library"
    assert not re.search(r"\bHMMA\.", sass) and "WGMMA" not in sass


def test_mma_issue_loops_stay_on_the_uniform_datapath():
    """a tcgen05.mma issued under `if (lane == 0)` compiles into an ELECT / R2UR.BROADCAST / branch loop
    per instruction (~20 SASS instructions between two UTCHMMA), and that paced every GEMM of the encoder.  With
    warp-uniform operands and one elect.sync per k-block the four MMAs are back to back.  Guard it in the SASS of the
    built library: in every GEMM / search / attention kernel the typical distance between consecutive UTCHMMA is a few
    instructions, and no R2UR.BROADCAST sits between two of them inside a k-block."""
    import shutil
    import subprocess

    from clip_lora_match_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    _lib.load()
    sass = subprocess.run([cuobjdump, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    checked = 0
    for fn in re.split(r"\n\s*Function : ", sass)[1:]:
        name = fn.split("\n", 1)[0]
        ins = [l for l in fn.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
        idx = [i for i, l in enumerate(ins) if "UTCHMMA" in l]
        if len(idx) < 4:
            continue
        gaps = sorted(b - a for a, b in zip(idx, idx[1:]))
        median = gaps[len(gaps) // 2]
        assert median <= 6, f"{name}: {median} instructions between consecutive UTCHMMA (waterfall issue loop?)"
        if "gemm_kernel" in name or "search_kernel" in name:
            # the four MMAs of a k-block: nothing but uniform-datapath work in between
            between = "\n".join(ins[idx[0]:idx[3] + 1])
            assert "R2UR.BROADCAST" not in between and "ELECT" not in between, name
        checked += 1
    assert checked >= 10, f"only {checked} kernels with tcgen05.mma found"


def test_extra_token_attention_decomposition_is_exact_softmax_attention():
    """The ViT-L/14 attention plan (csrc/clm_attention.cu, T = 128k + 1) restated in torch: tensor cores take the
    256 x 256 block (P rounded to bf16), the class token's key enters each row through one dot product (row max,
    row sum, rank-1 update of O in fp32) and its query row is a separate fp32 row.  Same result as plain softmax
    attention to bf16-P accuracy: the decomposition itself loses nothing."""
    g = torch.Generator().manual_seed(0)
    T, d = 257, 64
    q, k, v = (torch.randn((T, d), generator=g).bfloat16().float() * 1.5 for _ in range(3))
    ref = torch.softmax((q @ k.T) * 0.125, dim=-1) @ v
    Tk = T - 1
    c = 0.125 * 1.4426950408889634
    s_main = q[:Tk] @ k[:Tk].T                       # S = Q K^T on the tensor cores
    s_x = q[:Tk] @ k[Tk]                             # the extra key: one dot product per row
    m = torch.maximum(s_main.max(dim=-1).values, s_x)
    p_main = torch.exp2((s_main - m[:, None]) * c)
    p_x = torch.exp2((s_x - m) * c)
    denom = p_main.sum(dim=-1) + p_x                 # the row sum is taken in fp32, before P is rounded
    o = p_main.bfloat16().float() @ v[:Tk] + p_x[:, None] * v[Tk][None, :]   # P V (bf16 P) + rank-1 update
    out_main = o / denom[:, None]
    s_t = q[Tk] @ k.T                                # the extra query row, fp32 throughout
    p_t = torch.exp2((s_t - s_t.max()) * c)
    out_tail = (p_t @ v) / p_t.sum()
    out = torch.cat([out_main, out_tail[None, :]], dim=0)
    assert torch.allclose(out, ref, atol=6e-3, rtol=0)
    assert torch.allclose(out[Tk], ref[Tk], atol=1e-5)   # the tail row has no bf16 rounding at all


def test_pooling_eos_id_maps_the_legacy_value():
    """text_config.eos_token_id == 2 (every openai/clip-vit-* checkpoint) makes transformers pool at
    input_ids.argmax (TF:575-590), i.e. at the end-of-text token vocab-1; any other value pools at its first
    occurrence.  The tower's eos_id must follow."""
    from transformers import CLIPConfig

    from clip_lora_match_b200.models import clip_model as CM

    assert CM.pooling_eos_id(2, 49408) == 49407
    assert CM.pooling_eos_id(None, 49408) == 49407
    assert CM.pooling_eos_id(49407, 49408) == 49407
    assert CM.pooling_eos_id(1234, 49408) == 1234
    cfg = CLIPConfig(text_config={"eos_token_id": 2})  # what the published openai/clip-vit-* configs carry
    assert cfg.text_config.eos_token_id == 2
    assert CM.arch_from_hf_config(cfg).eos_id == cfg.text_config.vocab_size - 1


def test_train_lora_mirror_host_logic(tmp_path):
    """scripts/train_lora.py mirror: the schedule equals the reference's closure (golden from its own source), the
    config / CSV errors are the reference's, and the loss front end refuses CPU tensors (no fallback)."""
    import numpy as np

    from clip_lora_match_b200.scripts import train_lora as TL

    g = np.load(os.path.join(ROOT, "tests", "golden", "train_golden.npz"))
    for row in g["sched"]:
        total, warm = int(row[0]), int(row[1])
        assert [TL.lr_lambda(s, total, warm) for s in range(total + 2)] == list(row[2:2 + total + 2])
    with pytest.raises(FileNotFoundError, match="LoRA config file not found"):
        TL.load_lora_training_config(tmp_path / "nope.yaml")
    with pytest.raises(ValueError, match="train_csv"):
        TL.build_dataloaders({"data": {}})
    with pytest.raises(FileNotFoundError, match="CSV not found"):
        TL.ClipPairDataset(tmp_path / "nope.csv", processor=object())
    bad = tmp_path / "bad.csv"
    bad.write_text("a,b\n1,2\n")
    with pytest.raises(ValueError, match="image_path"):
        TL.ClipPairDataset(bad, processor=object())
    with pytest.raises(ValueError, match="CUDA"):
        TL.compute_clip_contrastive_loss(torch.randn(4, 8), torch.randn(4, 8))
    # the shipped YAML carries the reference's training section
    cfg = TL.load_lora_training_config(os.path.join(ROOT, "clip_lora_match_b200", "config", "lora_config.yaml"))
    assert cfg["training"]["temperature"] == 0.07 and cfg["training"]["batch_size"] == 8


def test_layernorm_fold_algebra_on_the_host():
    """kernels.fold_layernorm prepares what clm_gemm_ln_epi consumes (include/clm_b200.h): with Wg = W diag(gamma)
    rounded to bf16, col_sums of the ROUNDED values and bias' = b + W beta, the epilogue formula
    rstd (h Wg^T - mean col_sums) + bias' is LayerNorm(h) W^T + b up to the rounding of Wg -- and exactly the
    LayerNorm without affine part applied to Wg.  Same for the LoRA down-projection in its 1 / rstd form, whose constant
    (A beta)(sB)^T moves into the wide GEMM's bias."""
    from clip_lora_match_b200 import kernels as K

    g = torch.Generator().manual_seed(0)
    D, N, r = 256, 96, 16
    h = (torch.randn((40, D), generator=g) * 1.5 + torch.randn((40, 1), generator=g) * 3.0).bfloat16().float()
    w = torch.randn((N, D), generator=g) * D ** -0.5
    gamma = torch.rand((D,), generator=g) + 0.5
    beta = torch.randn((D,), generator=g) * 0.1
    bias = torch.randn((N,), generator=g)
    wg, cs, bf = K.fold_layernorm(w, gamma, beta, bias)
    assert wg.dtype == torch.bfloat16 and cs.dtype == torch.float32 and bf.dtype == torch.float32
    assert torch.equal(cs, wg.float().sum(dim=1))
    mu = h.mean(dim=1, keepdim=True)
    rstd = torch.rsqrt(h.var(dim=1, unbiased=False, keepdim=True) + 1e-5)
    folded = rstd * (h @ wg.float().T - mu * cs) + bf
    plain_ln = torch.nn.functional.layer_norm(h, (D,), None, None, 1e-5)
    assert torch.allclose(folded, plain_ln @ wg.float().T + bf, atol=2e-4, rtol=1e-4)          # exact in Wg
    ref = torch.nn.functional.layer_norm(h, (D,), gamma, beta, 1e-5) @ w.T + bias
    assert torch.allclose(folded, ref, atol=3e-2, rtol=1e-2)                                    # bf16 rounding of Wg
    # LoRA: u = (LN(h) A^T - A beta) / rstd from ln_mode 2 (no bias); t = rstd u + A beta
    a = torch.randn((r, D), generator=g) * D ** -0.5
    sb = torch.randn((N, r), generator=g) * 0.1
    ag, s_a, c_a = K.fold_layernorm(a, gamma, beta)
    u = h @ ag.float().T - mu * s_a
    t_ref = torch.nn.functional.layer_norm(h, (D,), gamma, beta, 1e-5) @ a.T
    assert torch.allclose(rstd * u + c_a, t_ref, atol=3e-2, rtol=1e-2)
    # the wide GEMM: acc = h Wg^T + u (sB)^T, epilogue rstd (acc - mean col_sums) + (bias' + (sB)(A beta))
    out = rstd * (h @ wg.float().T + u @ sb.T - mu * cs) + (bf + sb @ c_a)
    assert torch.allclose(out, ref + t_ref @ sb.T, atol=4e-2, rtol=1e-2)


def test_residual_stream_option_names():
    from clip_lora_match_b200.models import clip_model as CM

    assert CM.DEFAULT_RESIDUAL_DTYPE == "bfloat16"
    assert CM._residual_dtype_name("bf16") == "bfloat16" and CM._residual_dtype_name("fp32") == "float32"
    assert CM._residual_dtype_name(torch.bfloat16) == "bfloat16" and CM._residual_dtype_name(torch.float32) == "float32"
    with pytest.raises(ValueError):
        CM._residual_dtype_name("float16")
    os.environ["CLM_RESIDUAL_DTYPE"] = "float32"
    try:
        assert CM._residual_dtype_name(None) == "float32"
    finally:
        del os.environ["CLM_RESIDUAL_DTYPE"]
    assert CM._residual_dtype_name(None) == CM.DEFAULT_RESIDUAL_DTYPE
