"""CPU oracle for the retrieval hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under clip_lora_match_b200/ does.

It restates, in plain fp32 PyTorch on the CPU, the algorithm the reference
(youngalip/clip-lora-match) runs for this path.  The arithmetic of the encoder lives in
third-party packages that are NOT vendored in /root/reference:

  * transformers (unpinned in the reference's requirements.txt:4; 5.5.0 installed here) —
    CLIPModel, modeling_clip.py ("TF:" below).  The oracle instantiates the real
    transformers CLIPModel (eager attention, fp32, random init) so the encoder math is the
    library's own, not a re-derivation.
  * peft (unpinned, requirements.txt:6; NOT installed, no network) — restated in
    `LoraLinear` / `inject_lora` below from PEFT's documented LoRA semantics
    (SURVEY.md Appendix B), anchored on the reference's call sites models/lora_adapter.py:35-42,53
    and models/clip_model.py:78.

Parity pinning: the reference has no tests.  The search half of this oracle is
pinned against outputs of the reference's own src/embedding/similarity.py and
src/embedding/search.py run in the build container (oracle/make_golden.py ->
tests/golden/search_golden.npz), including the one fixture the reference ships
(data/index/custom_items_index.pt).  The encoder half is pinned against the reference's
own encode_image/encode_text/attach_lora_to_clip code executed over transformers 5.5 through a
4.x-compat shim and the peft stub (tests/golden/encoder_golden.npz); since the reference
publishes no encoder vectors and peft itself is absent, LoRA arithmetic is "parity unpinned"
against real peft (stated in DESIGN.md).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn as nn

# ------------------------------------------------------------------------------------------
# architectures (transformers/models/clip/configuration_clip.py defaults)
# ------------------------------------------------------------------------------------------
ARCH_TABLE = {
    # name: (v_width, v_layers, v_heads, v_mlp, patch, t_width, t_layers, t_heads, t_mlp, proj)
    "openai/clip-vit-base-patch32": (768, 12, 12, 3072, 32, 512, 12, 8, 2048, 512),
    "openai/clip-vit-base-patch16": (768, 12, 12, 3072, 16, 512, 12, 8, 2048, 512),
    "openai/clip-vit-large-patch14": (1024, 24, 16, 4096, 14, 768, 12, 12, 3072, 768),
    # reduced-size configuration for fast CPU tests (same code paths, head_dim 64)
    "tiny-test": (128, 2, 2, 256, 32, 128, 2, 2, 256, 64),
}
BOS_ID, EOS_ID = 49406, 49407


def hf_config(name: str):
    from transformers import CLIPConfig

    vw, vl, vh, vm, p, tw, tl, th, tm, proj = ARCH_TABLE[name]
    return CLIPConfig(
        text_config=dict(hidden_size=tw, num_hidden_layers=tl, num_attention_heads=th,
                         intermediate_size=tm, max_position_embeddings=77, vocab_size=49408,
                         projection_dim=proj, bos_token_id=BOS_ID, eos_token_id=EOS_ID,
                         hidden_act="quick_gelu", layer_norm_eps=1e-5),
        vision_config=dict(hidden_size=vw, num_hidden_layers=vl, num_attention_heads=vh,
                           intermediate_size=vm, patch_size=p, image_size=224, projection_dim=proj,
                           hidden_act="quick_gelu", layer_norm_eps=1e-5),
        projection_dim=proj)


def build_model(name: str, seed: int = 0):
    """Random-init transformers CLIPModel (HF init scheme TF:403-459), fp32, eager attention."""
    from transformers import CLIPModel

    cfg = hf_config(name)
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    model = CLIPModel(cfg)
    model.eval()
    return model.float()


# ------------------------------------------------------------------------------------------
# PEFT LoRA semantics (Appendix B)
# ------------------------------------------------------------------------------------------
class LoraLinear(nn.Module):
    """y = base(x) + lora_B(lora_A(dropout(x))) * (alpha / r); dropout is identity in eval()."""

    def __init__(self, base: nn.Linear, r: int, alpha: float, dropout: float = 0.0):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.Linear(base.in_features, r, bias=False)
        self.lora_B = nn.Linear(r, base.out_features, bias=False)
        self.lora_dropout = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        self.scaling = alpha / r
        nn.init.kaiming_uniform_(self.lora_A.weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B.weight)
        for p in base.parameters():
            p.requires_grad_(False)

    def forward(self, x):
        return self.base_layer(x) + self.lora_B(self.lora_A(self.lora_dropout(x))) * self.scaling


def target_paths(model: nn.Module, target_modules: Iterable[str]) -> List[str]:
    """PEFT list-matching: name == t or name endswith '.'+t, over nn.Linear modules."""
    targets = list(target_modules)
    out = []
    for name, mod in model.named_modules():
        if isinstance(mod, nn.Linear) and any(name == t or name.endswith("." + t) for t in targets):
            out.append(name)
    return out


def inject_lora(model: nn.Module, r: int, alpha: float, target_modules: Iterable[str],
                dropout: float = 0.0) -> List[str]:
    """Wrap every matching Linear (both towers) in place; returns the wrapped paths."""
    paths = target_paths(model, target_modules)
    for path in paths:
        parent_name, _, leaf = path.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        setattr(parent, leaf, LoraLinear(getattr(parent, leaf), r, alpha, dropout))
    return paths


def set_lora_weights(model: nn.Module, weights: Dict[str, Tuple[torch.Tensor, torch.Tensor]]) -> None:
    for path, (a, b) in weights.items():
        mod = model.get_submodule(path)
        assert isinstance(mod, LoraLinear), path
        with torch.no_grad():
            mod.lora_A.weight.copy_(a)
            mod.lora_B.weight.copy_(b)


def get_lora_weights(model: nn.Module) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    return {n: (m.lora_A.weight.detach().clone(), m.lora_B.weight.detach().clone())
            for n, m in model.named_modules() if isinstance(m, LoraLinear)}


def base_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    """State dict with HF CLIPModel names (LoRA wrappers unwrapped: '.base_layer.' removed)."""
    out = {}
    for k, v in model.state_dict().items():
        if ".lora_A." in k or ".lora_B." in k:
            continue
        out[k.replace(".base_layer.", ".")] = v.detach().clone()
    return out


def synthetic_lora(model: nn.Module, r: int, alpha: float, target_modules: Iterable[str],
                   seed: int = 1, b_std: float = 0.02) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    """A kaiming-uniform (PEFT default), B ~ N(0, 0.02^2) so LoRA is NOT a no-op."""
    paths = inject_lora(model, r, alpha, target_modules)
    g = torch.Generator().manual_seed(seed)
    weights = {}
    for p in paths:
        mod = model.get_submodule(p)
        a = torch.empty_like(mod.lora_A.weight)
        bound = 1.0 / math.sqrt(a.shape[1])
        a.uniform_(-bound, bound, generator=g)
        b = torch.randn(mod.lora_B.weight.shape, generator=g) * b_std
        weights[p] = (a, b)
    set_lora_weights(model, weights)
    return weights


# ------------------------------------------------------------------------------------------
# encoder: exactly the reference's post-processing around the library forward
# ------------------------------------------------------------------------------------------
def _pooled(out):
    # transformers 5.x returns BaseModelOutputWithPooling (TF:829-863); 4.x returned the tensor
    return out.pooler_output if hasattr(out, "pooler_output") else out


@torch.no_grad()
def encode_images(model, pixel_values: torch.Tensor, normalize: bool = True, batch_size: int = 16) -> torch.Tensor:
    """reference models/clip_model.py:114-116 (and embed_image.py:87-91), batched."""
    outs = []
    for i in range(0, pixel_values.shape[0], batch_size):
        f = _pooled(model.get_image_features(pixel_values=pixel_values[i:i + batch_size].float()))
        if normalize:
            f = f / f.norm(dim=-1, keepdim=True)
        outs.append(f)
    return torch.cat(outs, 0) if outs else torch.empty(0)


@torch.no_grad()
def encode_texts(model, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                 normalize: bool = True, batch_size: int = 16) -> torch.Tensor:
    """reference models/clip_model.py:143-148 (and embed_text.py:46-53), batched."""
    outs = []
    for i in range(0, input_ids.shape[0], batch_size):
        ids = input_ids[i:i + batch_size].long()
        am = attention_mask[i:i + batch_size] if attention_mask is not None else None
        f = _pooled(model.get_text_features(input_ids=ids, attention_mask=am))
        if normalize:
            f = f / f.norm(dim=-1, keepdim=True)
        outs.append(f)
    return torch.cat(outs, 0) if outs else torch.empty(0)


# ------------------------------------------------------------------------------------------
# search: reference src/embedding/search.py and similarity.py restated
# ------------------------------------------------------------------------------------------
def normalize_rows(x: torch.Tensor) -> torch.Tensor:
    """x / x.norm(dim=-1, keepdim=True) — no epsilon (search.py:68,93; similarity.py:28-29)."""
    return x / x.norm(dim=-1, keepdim=True)


def search_topk(index_embeddings: torch.Tensor, queries: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """search.py:68 (row-normalise index), :93 (normalise query), :96 matmul, :98 k=min(k,N),
    :99 topk(largest, sorted) — batched over queries [Q, d]."""
    e = normalize_rows(index_embeddings.float())
    q = queries.float()
    if q.dim() == 1:
        q = q.unsqueeze(0)
    q = normalize_rows(q)
    sims = q @ e.T
    k = min(k, e.shape[0])
    return torch.topk(sims, k=k, dim=-1, largest=True, sorted=True)


def fuse_query(text_emb: Optional[torch.Tensor], image_emb: Optional[torch.Tensor],
               w_text: float = 0.5, w_image: float = 0.5) -> torch.Tensor:
    """reference src/embedding/seeker_service.py:139-157 restated: one modality -> renormalise;
    two -> python `sum(w * e ...)` (0 + w_t*e_t + w_i*e_i) -> renormalise.  Works on (d,) or [Q,d]."""
    embs = []
    if text_emb is not None:
        embs.append((text_emb.float(), w_text))
    if image_emb is not None:
        embs.append((image_emb.float(), w_image))
    if not embs:
        raise ValueError("Minimal harus ada query_text atau query_image_path.")
    if len(embs) == 1:
        emb, _ = embs[0]
        return emb / emb.norm(dim=-1, keepdim=True)
    weighted = sum(w * e for e, w in embs)
    return weighted / weighted.norm(dim=-1, keepdim=True)


def ids_match_with_ties(ref_scores: torch.Tensor, ref_ids: torch.Tensor, got_ids: torch.Tensor,
                        sims: torch.Tensor, tol: float = 1e-4) -> bool:
    """north_star's rule: ids identical except where the oracle scores tie within `tol`."""
    mism = ref_ids != got_ids
    if not mism.any():
        return True
    s_got = torch.gather(sims, 1, got_ids)
    return bool(((s_got - ref_scores).abs()[mism] <= tol).all())


# ------------------------------------------------------------------------------------------
# synthetic inputs (seeds fixed)
# ------------------------------------------------------------------------------------------
def synth_images(batch: int, seed: int = 2, image: int = 224) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn((batch, 3, image, image), generator=g)


def synth_captions(batch: int, seed: int = 3, context: int = 77) -> Tuple[torch.Tensor, torch.Tensor]:
    """ids [B,77] = [BOS, tok..., EOS, EOS...]; tok ~ U[0,49405]; length L ~ U{3..77}; mask = pos < L."""
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(3, context + 1, (batch,), generator=g)
    ids = torch.randint(0, BOS_ID, (batch, context), generator=g)
    pos = torch.arange(context).unsqueeze(0)
    ids[:, 0] = BOS_ID
    ids = torch.where(pos >= (lengths - 1).unsqueeze(1), torch.full_like(ids, EOS_ID), ids)
    mask = (pos < lengths.unsqueeze(1)).long()
    return ids, mask


def synth_unit_rows(n: int, d: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return normalize_rows(torch.randn((n, d), generator=g))


# ------------------------------------------------------------------------------------------
# parity metrics (plain cosine is weak on random-init weights)
# ------------------------------------------------------------------------------------------
def parity_metrics(got: torch.Tensor, ref: torch.Tensor) -> Dict[str, float]:
    got, ref = got.float().cpu(), ref.float().cpu()
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=-1)
    mean = ref.mean(dim=0, keepdim=True)
    ccos = torch.nn.functional.cosine_similarity(got - mean, ref - mean, dim=-1)
    rel = (got - ref).norm(dim=-1) / ref.norm(dim=-1)
    return {"cos_min": float(cos.min()), "centered_cos_min": float(ccos.min()),
            "rel_l2_max": float(rel.max())}
