"""CLIP + LoRA encoder façade — the B200 mirror of the reference's models/clip_model.py.

Reference surface kept (names, positional order, return types, exceptions):
  load_clip_model(config_path, use_lora, lora_weights_path) -> (model, processor, device)
                                                  reference models/clip_model.py:37-82
  encode_image(image_path, model, processor, device) -> (d,) CPU fp32   reference :89-118
  encode_text(text, model, processor, device)        -> (d,) CPU fp32   reference :121-150
Batched extensions (the reference is batch-1 everywhere):
  B200ClipModel.encode_images(pixel_values) / encode_texts(input_ids) -> [B, d] device fp32.

`model` is a B200ClipModel: it owns bf16/fp32 device copies of the weights and two
clm_tower handles (include/clm_b200.h); every FLOP of the forward runs in the sm_100a
kernels of csrc/.  transformers is used only as a weight container (from_pretrained /
random init of the named architecture) and for the host-side image processor.
"""
from __future__ import annotations

import collections
import ctypes as C
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Tuple, Union

import torch
import yaml

from .. import _lib
from .. import kernels as K
from .lora_adapter import LoraAdapter, linear_module_paths, load_lora_adapter

LORA_COLS = 64  # LoRA ranks are padded to multiples of this: one extra K block of the fused GEMMs each
BOS_ID, EOS_ID = 49406, 49407


# ------------------------------------------------------------------------------------------
# architectures
# ------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class TowerArch:
    width: int
    layers: int
    heads: int
    mlp: int


@dataclass(frozen=True)
class ClipArch:
    name: str
    vision: TowerArch
    text: TowerArch
    patch: int
    proj_dim: int
    image: int = 224
    context: int = 77
    vocab: int = 49408
    eos_id: int = EOS_ID
    ln_eps: float = 1e-5

    @property
    def vision_tokens(self) -> int:
        return 1 + (self.image // self.patch) ** 2


ARCHS: Dict[str, ClipArch] = {
    "openai/clip-vit-base-patch32": ClipArch(
        "openai/clip-vit-base-patch32", TowerArch(768, 12, 12, 3072), TowerArch(512, 12, 8, 2048), 32, 512),
    "openai/clip-vit-base-patch16": ClipArch(
        "openai/clip-vit-base-patch16", TowerArch(768, 12, 12, 3072), TowerArch(512, 12, 8, 2048), 16, 512),
    "openai/clip-vit-large-patch14": ClipArch(
        "openai/clip-vit-large-patch14", TowerArch(1024, 24, 16, 4096), TowerArch(768, 12, 12, 3072), 14, 768),
}


def arch_from_name(name: str) -> ClipArch:
    if name in ARCHS:
        return ARCHS[name]
    raise ValueError(f"unknown CLIP architecture {name!r}; known: {sorted(ARCHS)}")


def pooling_eos_id(eos_token_id, vocab_size: int) -> int:
    """The token id whose FIRST occurrence the text tower pools at.

    transformers pools at `input_ids.argmax(-1)` when `text_config.eos_token_id == 2` (the legacy value
    every openai/clip-vit-* checkpoint still carries; TF:575-590) -- i.e. at the highest id of the row,
    which for CLIP's tokenizer is the end-of-text token `vocab_size - 1` (49407) -- and at the first
    occurrence of `eos_token_id` otherwise.  Both rules are "first position of the EOS token" once the
    legacy value (or a missing one) is mapped to `vocab_size - 1`."""
    if eos_token_id is None or int(eos_token_id) == 2:
        return int(vocab_size) - 1
    return int(eos_token_id)


def arch_from_hf_config(cfg, name: str = "custom") -> ClipArch:
    v, t = cfg.vision_config, cfg.text_config
    return ClipArch(
        name=name,
        vision=TowerArch(v.hidden_size, v.num_hidden_layers, v.num_attention_heads, v.intermediate_size),
        text=TowerArch(t.hidden_size, t.num_hidden_layers, t.num_attention_heads, t.intermediate_size),
        patch=v.patch_size, proj_dim=cfg.projection_dim, image=v.image_size,
        context=t.max_position_embeddings, vocab=t.vocab_size,
        eos_id=pooling_eos_id(t.eos_token_id, t.vocab_size),
        ln_eps=v.layer_norm_eps)


def hf_config_for(arch: ClipArch):
    """transformers CLIPConfig of the architecture (random-init container and oracle use it)."""
    from transformers import CLIPConfig

    return CLIPConfig(
        text_config=dict(hidden_size=arch.text.width, num_hidden_layers=arch.text.layers,
                         num_attention_heads=arch.text.heads, intermediate_size=arch.text.mlp,
                         max_position_embeddings=arch.context, vocab_size=arch.vocab,
                         projection_dim=arch.proj_dim, bos_token_id=BOS_ID, eos_token_id=arch.eos_id,
                         hidden_act="quick_gelu", layer_norm_eps=arch.ln_eps),
        vision_config=dict(hidden_size=arch.vision.width, num_hidden_layers=arch.vision.layers,
                           num_attention_heads=arch.vision.heads, intermediate_size=arch.vision.mlp,
                           patch_size=arch.patch, image_size=arch.image, projection_dim=arch.proj_dim,
                           hidden_act="quick_gelu", layer_norm_eps=arch.ln_eps),
        projection_dim=arch.proj_dim)


# ------------------------------------------------------------------------------------------
# the model object
# ------------------------------------------------------------------------------------------
class _GraphEntry:
    """A captured tower pass: the graph, the output buffer its kernels write, its launch count."""
    __slots__ = ("graph", "inp", "out", "launches")

    def __init__(self):
        self.graph = None
        self.inp = None
        self.out = None
        self.launches = 0


# Residual stream between the transformer layers.  The reference runs fp32 end to end; "bfloat16" keeps the stream
# in bf16 (one extra rounding per residual update; embeddings measured at cosine >= 0.9999 of the fp32 stream and
# >= 0.999 -- north_star's bar -- of the fp32 CPU oracle on every architecture, tests/test_residual_bf16_gpu.py).
# bf16 is the default: BASELINE's configs name bf16 as the compute type, and the stream's type is worth 4-5 % of the
# ViT-L/14 step (profiles/r2_residual_ab.log).  "float32" (argument, or CLM_RESIDUAL_DTYPE=float32) is the
# reference's stream.
DEFAULT_RESIDUAL_DTYPE = "bfloat16"
_RESIDUAL_CODES = {"float32": _lib.OUT_F32, "bfloat16": _lib.OUT_BF16}
_RESIDUAL_ALIASES = {"f32": "float32", "fp32": "float32", "float": "float32", "bf16": "bfloat16"}


def _residual_dtype_name(name: Optional[str]) -> str:
    if name is None:
        name = os.environ.get("CLM_RESIDUAL_DTYPE") or DEFAULT_RESIDUAL_DTYPE
    name = str(name).replace("torch.", "").lower()
    name = _RESIDUAL_ALIASES.get(name, name)
    if name not in _RESIDUAL_CODES:
        raise ValueError(f"residual_dtype must be 'float32' or 'bfloat16', not {name!r}")
    return name


class B200ClipModel:
    """CLIP dual encoder whose forward is the C-ABI of include/clm_b200.h.

    Quacks like the object the reference passes around (`.eval()`, `.to()`, `.parameters()`,
    `.get_image_features`, `.get_text_features`) so reference-style callers keep working.
    """

    def __init__(self, arch: ClipArch, state_dict: Dict[str, torch.Tensor],
                 lora: Optional[LoraAdapter] = None, device: Union[str, torch.device] = "cuda",
                 max_workspace_bytes: int = 24 << 30, residual_dtype: Optional[str] = None,
                 ln_fold: Optional[bool] = None):
        """residual_dtype: "float32" (the reference's fp32 residual stream) or "bfloat16" (the stream between the
        layers is stored in bf16: clm_tower_set_residual_dtype in include/clm_b200.h); None takes
        DEFAULT_RESIDUAL_DTYPE (overridable with CLM_RESIDUAL_DTYPE).
        ln_fold: with a bf16 stream, fold layer_norm1 / layer_norm2 into the QKV / fc1 GEMMs (gamma into the weights,
        mean / rstd applied in the epilogue: clm_gemm_ln_epi) instead of running LayerNorm passes; None = on unless
        CLM_LN_FOLD=0."""
        self.arch = arch
        self.name = arch.name
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("B200ClipModel needs a CUDA device: the sm_100a path has no CPU fallback")
        self._lib = _lib.load()
        _lib.check(self._lib.clm_device_check(), "clm_device_check")
        self._sd = {k: v.detach().to("cpu", torch.float32) for k, v in state_dict.items()}
        self.lora: Optional[LoraAdapter] = None
        self.max_workspace_bytes = max_workspace_bytes
        self._towers: Dict[str, int] = {}
        self._keep: Dict[str, list] = {}
        self._folds: Dict[str, object] = {}
        # LayerNorm folded into the QKV / fc1 GEMMs (used by the towers only while the residual stream is bf16)
        self.ln_fold = os.environ.get("CLM_LN_FOLD", "1") != "0" if ln_fold is None else bool(ln_fold)
        self._workspace: Optional[torch.Tensor] = None
        self._dummy = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._stage = None  # host->device staging (encode_images of host tensors)
        # CUDA graphs of whole tower passes for small batches, keyed by (tower, batch, ...): see _run_tower.
        # CLM_GRAPHS=0 disables.
        self.use_graphs = os.environ.get("CLM_GRAPHS", "1") != "0"
        self._graphs: "collections.OrderedDict[tuple, _GraphEntry]" = collections.OrderedDict()
        self.residual_dtype = _residual_dtype_name(residual_dtype)
        self.set_lora(lora)

    # ---- reference-compat surface ------------------------------------------------------
    def eval(self):
        return self

    def to(self, *args, **kwargs):
        return self

    def parameters(self):
        # the reference asks `next(model.parameters()).dtype` to cast pixel_values
        # (models/clip_model.py:111-112); the C-ABI takes fp32 pixel_values.
        yield self._dummy

    def num_base_parameters(self) -> int:
        return sum(v.numel() for v in self._sd.values())

    def linear_dims(self) -> Dict[str, Tuple[int, int]]:
        dims = {}
        for p in linear_module_paths(self.arch.vision.layers, self.arch.text.layers):
            w = self._sd.get(p + ".weight")
            if w is not None:
                dims[p] = (w.shape[0], w.shape[1])
        return dims

    # ---- weights -----------------------------------------------------------------------
    def set_lora(self, lora: Optional[LoraAdapter]) -> None:
        """(Re)build both towers with the given unmerged adapter (None = base model)."""
        self._destroy_towers()
        self.lora = lora
        for kind in ("vision", "text"):
            self._build_tower(kind)

    def set_residual_dtype(self, residual_dtype: Optional[str]) -> None:
        """Switch the residual stream of both towers between "float32" and "bfloat16" (weights untouched)."""
        self.residual_dtype = _residual_dtype_name(residual_dtype)
        self._graphs.clear()  # captured passes carry the stream's type
        for kind, h in self._towers.items():
            _lib.check(self._lib.clm_tower_set_residual_dtype(h, _RESIDUAL_CODES[self.residual_dtype]),
                       f"clm_tower_set_residual_dtype({kind})")

    def set_ln_fold(self, on: bool) -> None:
        """Use (or stop using) the folded-LayerNorm GEMMs; needs the folded weights built at construction."""
        self._graphs.clear()
        for kind, h in self._towers.items():
            folds = self._folds.get(kind) if on else None
            if on and folds is None:
                raise ValueError("the model was built with ln_fold=False: no folded weights exist")
            _lib.check(self._lib.clm_tower_set_ln_fold(h, folds), f"clm_tower_set_ln_fold({kind})")

    def _dev(self, t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        return t.to(device=self.device, dtype=dtype).contiguous()

    def _lora_operands(self, prefix: str, in_dim: int, group: List[Tuple[str, int]]):
        """Pack the adapters of `group` -- (module leaf, out features) of the Linears that share one fused
        GEMM and one input -- into A_cat [cols, in_dim] and (s*B)_cat [sum(out), cols], cols = the total
        rank padded to a multiple of LORA_COLS (one K block of the GEMM).  (None, None, 0) if no member has
        LoRA.  Any rank and any subset of the group is accepted (reference models/lora_adapter.py:33-41)."""
        if self.lora is None:
            return None, None, 0
        present = [(i, m) for i, (m, _) in enumerate(group) if f"{prefix}.{m}" in self.lora.weights]
        if not present:
            return None, None, 0
        r = self.lora.config.r
        cols = (r * len(present) + LORA_COLS - 1) // LORA_COLS * LORA_COLS
        outs = [o for _, o in group]
        a_cat = torch.zeros((cols, in_dim), dtype=torch.float32)
        b_cat = torch.zeros((sum(outs), cols), dtype=torch.float32)
        s = self.lora.scaling
        for slot, (i, m) in enumerate(present):
            a, b = self.lora.weights[f"{prefix}.{m}"]
            if tuple(a.shape) != (r, in_dim) or tuple(b.shape) != (outs[i], r):
                raise ValueError(f"LoRA shape mismatch on {prefix}.{m}: A {tuple(a.shape)} B {tuple(b.shape)} "
                                 f"(expected A [{r}, {in_dim}], B [{outs[i]}, {r}])")
            row0 = sum(outs[:i])
            a_cat[slot * r:(slot + 1) * r] = a
            b_cat[row0:row0 + outs[i], slot * r:(slot + 1) * r] = b * s
        self._last_a_cat_f32 = a_cat  # fp32 host copy: the LayerNorm fold scales it by gamma before rounding
        return self._dev(a_cat, torch.bfloat16), self._dev(b_cat, torch.bfloat16), cols

    def _build_tower(self, kind: str) -> None:
        arch = self.arch
        ta = arch.vision if kind == "vision" else arch.text
        pre = "vision_model" if kind == "vision" else "text_model"
        sd = self._sd
        keep: list = []

        def f32(name):
            t = self._dev(sd[name], torch.float32)
            keep.append(t)
            return t.data_ptr()

        def bf16(t):
            t = self._dev(t, torch.bfloat16)
            keep.append(t)
            return t.data_ptr()

        layers = (_lib.LayerWeights * ta.layers)()
        folds = (_lib.LayerLnFold * ta.layers)()
        cols = {"qkv": 0, "out": 0, "fc1": 0, "fc2": 0}  # every layer of a tower carries the same adapter shape
        a_f32: Dict[str, torch.Tensor] = {}   # fp32 host A_cat of this layer's adapters, per fused GEMM
        b_bf16: Dict[str, torch.Tensor] = {}  # the (s B)_cat the device multiplies with

        def fold(Fl, name, w, gamma, beta, bias, names):
            """LayerNorm folded into the Linear that consumes it (include/clm_b200.h: clm_gemm_ln_epi)."""
            wg, cs, bf = K.fold_layernorm(w, gamma, beta, bias)
            for field, t, dt in zip(names, (wg, cs, bf), (torch.bfloat16, torch.float32, torch.float32)):
                if field is None:
                    continue
                t = self._dev(t, dt)
                keep.append(t)
                setattr(Fl, field, t.data_ptr())

        def lora(L, name, prefix, in_dim, group):
            a_cat, b_cat, c = self._lora_operands(prefix, in_dim, group)
            if a_cat is None:
                return
            if cols[name] not in (0, c):
                raise ValueError(f"LoRA targets differ between layers of the {kind} tower ({name}: {cols[name]} vs {c})")
            cols[name] = c
            keep.extend([a_cat, b_cat])
            a_f32[name] = self._last_a_cat_f32
            b_bf16[name] = b_cat
            setattr(L, f"lora_a_{name if name != 'out' else 'o'}", a_cat.data_ptr())
            setattr(L, f"lora_b_{name if name != 'out' else 'o'}", b_cat.data_ptr())

        for i in range(ta.layers):
            lp = f"{pre}.encoder.layers.{i}"
            ap = f"{lp}.self_attn"
            L = layers[i]
            L.ln1_g, L.ln1_b = f32(f"{lp}.layer_norm1.weight"), f32(f"{lp}.layer_norm1.bias")
            L.ln2_g, L.ln2_b = f32(f"{lp}.layer_norm2.weight"), f32(f"{lp}.layer_norm2.bias")
            L.w_qkv = bf16(torch.cat([sd[f"{ap}.q_proj.weight"], sd[f"{ap}.k_proj.weight"],
                                      sd[f"{ap}.v_proj.weight"]], dim=0))
            bq = self._dev(torch.cat([sd[f"{ap}.q_proj.bias"], sd[f"{ap}.k_proj.bias"],
                                      sd[f"{ap}.v_proj.bias"]], dim=0), torch.float32)
            keep.append(bq)
            L.b_qkv = bq.data_ptr()
            lora(L, "qkv", ap, ta.width, [("q_proj", ta.width), ("k_proj", ta.width), ("v_proj", ta.width)])
            L.w_o, L.b_o = bf16(sd[f"{ap}.out_proj.weight"]), f32(f"{ap}.out_proj.bias")
            lora(L, "out", ap, ta.width, [("out_proj", ta.width)])
            L.w_fc1, L.b_fc1 = bf16(sd[f"{lp}.mlp.fc1.weight"]), f32(f"{lp}.mlp.fc1.bias")
            L.w_fc2, L.b_fc2 = bf16(sd[f"{lp}.mlp.fc2.weight"]), f32(f"{lp}.mlp.fc2.bias")
            lora(L, "fc1", f"{lp}.mlp", ta.width, [("fc1", ta.mlp)])
            lora(L, "fc2", f"{lp}.mlp", ta.mlp, [("fc2", ta.width)])
            if self.ln_fold:
                Fl = folds[i]
                g1, b1 = sd[f"{lp}.layer_norm1.weight"], sd[f"{lp}.layer_norm1.bias"]
                g2, b2 = sd[f"{lp}.layer_norm2.weight"], sd[f"{lp}.layer_norm2.bias"]
                # the adapter's share of the folded bias: LN(h) A^T = rstd u + A beta, so every row gets (A beta)(sB)^T
                extra_q = extra_1 = None
                if "qkv" in a_f32:
                    extra_q = b_bf16["qkv"].float().cpu() @ (a_f32["qkv"] @ b1.float())
                    fold(Fl, "a_qkv", a_f32["qkv"], g1, b1, None, ("lora_a_qkv_g", "s_a_qkv", None))
                if "fc1" in a_f32:
                    extra_1 = b_bf16["fc1"].float().cpu() @ (a_f32["fc1"] @ b2.float())
                    fold(Fl, "a_fc1", a_f32["fc1"], g2, b2, None, ("lora_a_fc1_g", "s_a_fc1", None))
                fold(Fl, "qkv", torch.cat([sd[f"{ap}.q_proj.weight"], sd[f"{ap}.k_proj.weight"],
                                           sd[f"{ap}.v_proj.weight"]], dim=0), g1, b1,
                     bq.cpu() if extra_q is None else bq.cpu() + extra_q, ("w_qkv_g", "s_qkv", "b_qkv_f"))
                fold(Fl, "fc1", sd[f"{lp}.mlp.fc1.weight"], g2, b2,
                     sd[f"{lp}.mlp.fc1.bias"] if extra_1 is None else sd[f"{lp}.mlp.fc1.bias"] + extra_1,
                     ("w_fc1_g", "s_fc1", "b_fc1_f"))
            a_f32.clear()
            b_bf16.clear()

        w = _lib.TowerWeights()
        cfg = _lib.TowerConfig()
        cfg.width, cfg.layers, cfg.heads, cfg.mlp = ta.width, ta.layers, ta.heads, ta.mlp
        cfg.proj_dim, cfg.ln_eps = arch.proj_dim, arch.ln_eps
        cfg.lora_cols_qkv, cfg.lora_cols_out = cols["qkv"], cols["out"]
        cfg.lora_cols_fc1, cfg.lora_cols_fc2 = cols["fc1"], cols["fc2"]
        w.pos_emb = f32(f"{pre}.embeddings.position_embedding.weight")
        if kind == "vision":
            cfg.kind, cfg.tokens, cfg.image, cfg.patch = 0, arch.vision_tokens, arch.image, arch.patch
            k = 3 * arch.patch * arch.patch
            kpad = (k + 63) // 64 * 64
            pw = torch.zeros((ta.width, kpad), dtype=torch.float32)
            pw[:, :k] = sd[f"{pre}.embeddings.patch_embedding.weight"].reshape(ta.width, k)
            w.patch_w = bf16(pw)
            w.class_emb = f32(f"{pre}.embeddings.class_embedding")
            w.pre_ln_g, w.pre_ln_b = f32(f"{pre}.pre_layrnorm.weight"), f32(f"{pre}.pre_layrnorm.bias")
            w.final_ln_g, w.final_ln_b = f32(f"{pre}.post_layernorm.weight"), f32(f"{pre}.post_layernorm.bias")
            w.proj_w = bf16(sd["visual_projection.weight"])
        else:
            cfg.kind, cfg.tokens, cfg.vocab, cfg.eos_id = 1, arch.context, arch.vocab, arch.eos_id
            w.tok_emb = f32(f"{pre}.embeddings.token_embedding.weight")
            w.final_ln_g, w.final_ln_b = f32(f"{pre}.final_layer_norm.weight"), f32(f"{pre}.final_layer_norm.bias")
            w.proj_w = bf16(sd["text_projection.weight"])
        handle = C.c_void_p()
        _lib.check(self._lib.clm_tower_create(C.byref(cfg), C.byref(w), layers, C.byref(handle)),
                   f"clm_tower_create({kind})")
        self._towers[kind] = handle.value
        self._keep[kind] = keep
        self._folds[kind] = folds if self.ln_fold else None
        if self.ln_fold:
            _lib.check(self._lib.clm_tower_set_ln_fold(handle.value, folds), f"clm_tower_set_ln_fold({kind})")
        _lib.check(self._lib.clm_tower_set_residual_dtype(handle.value, _RESIDUAL_CODES[self.residual_dtype]),
                   f"clm_tower_set_residual_dtype({kind})")

    def _destroy_towers(self) -> None:
        self._graphs.clear()  # captured launches point at the towers' weights
        for h in self._towers.values():
            self._lib.clm_tower_destroy(h)
        self._towers.clear()
        self._keep.clear()
        self._folds.clear()

    def __del__(self):
        try:
            self._destroy_towers()
        except Exception:
            pass

    # ---- forward -----------------------------------------------------------------------
    def _ensure_workspace(self, kind: str, batch: int) -> torch.Tensor:
        need = self._lib.clm_tower_workspace_bytes(self._towers[kind], batch)
        one = self._lib.clm_tower_workspace_bytes(self._towers[kind], 1)
        want = max(one, min(need, self.max_workspace_bytes))
        if self._workspace is None or self._workspace.numel() < want:
            self._workspace = None
            self._workspace = torch.empty(want, dtype=torch.uint8, device=self.device)
        return self._workspace

    H2D_CHUNK = 256  # images per host->device chunk when the input lives in host memory
    H2D_FIRST_CHUNK = 64  # ... except the first, whose copy is exposed
    MAX_GRAPHS = 16
    GRAPH_MAX_BATCH = 32

    def _run_tower(self, kind: str, inp: torch.Tensor, out: torch.Tensor, normalize: bool) -> None:
        """One tower pass over a device-resident batch: `inp` -> `out` on the current stream.

        Small batches (<= GRAPH_MAX_BATCH items: the reference's one-item-at-a-time calls) are launch
        bound -- ~100 kernels of a few microseconds each -- so their pass is captured once per batch size
        into a CUDA graph over graph-owned input / output buffers and replayed (first call eager, second
        call captures).  Large batches run eagerly: measured at batch 1024 the step is power bound and a
        graph changes nothing (tools/graph_probe.py, profiles/r1_notes.md)."""
        b = inp.shape[0]
        ws = self._ensure_workspace(kind, b)
        tower = self._towers[kind]
        tokens = inp.shape[1] if kind == "text" else 0  # a text pass may cover fewer positions than the context

        def launch(src: torch.Tensor, dst: torch.Tensor) -> None:
            if kind == "vision":
                rc = self._lib.clm_encode_image(tower, src.data_ptr(), b, dst.data_ptr(), int(normalize),
                                                ws.data_ptr(), ws.numel(), _lib.cur_stream())
            else:
                rc = self._lib.clm_encode_text_len(tower, src.data_ptr(), b, tokens, dst.data_ptr(), int(normalize),
                                                   ws.data_ptr(), ws.numel(), _lib.cur_stream())
            _lib.check(rc, f"clm_encode_{'image' if kind == 'vision' else 'text'}")

        if (not self.use_graphs or b > self.GRAPH_MAX_BATCH or self._lib.clm_prof_is_enabled()
                or torch.cuda.is_current_stream_capturing()):
            launch(inp, out)
            return
        key = (kind, b, tokens, bool(normalize), ws.data_ptr(), ws.numel())
        ent = self._graphs.get(key)
        if ent is None:
            launch(inp, out)  # also performs the kernels' one-time cudaFuncSetAttribute calls outside a capture
            self._graphs[key] = _GraphEntry()
            while len(self._graphs) > self.MAX_GRAPHS:
                self._graphs.popitem(last=False)
            return
        self._graphs.move_to_end(key)
        if ent.graph is None:
            ent.inp, ent.out = torch.empty_like(inp), torch.empty_like(out)
            n0 = self._lib.clm_launch_count()
            graph = torch.cuda.CUDAGraph()
            # thread_local: another thread (the NCCL watchdog) may touch the CUDA API during the capture
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                launch(ent.inp, ent.out)
            ent.launches = self._lib.clm_launch_count() - n0
            self._lib.clm_launch_count_add(-ent.launches)  # counted while capturing, but nothing has run yet
            ent.graph = graph
        ent.inp.copy_(inp)
        ent.graph.replay()
        self._lib.clm_launch_count_add(ent.launches)
        out.copy_(ent.out)

    def _encode_images_dev(self, pv: torch.Tensor, out: torch.Tensor, normalize: bool) -> None:
        self._run_tower("vision", pv, out, normalize)

    def encode_images(self, pixel_values: torch.Tensor, normalize: bool = True) -> torch.Tensor:
        """[B,3,H,W] fp32 (any device) -> [B, proj_dim] fp32 on the GPU; the batched form of
        reference encode_image (models/clip_model.py:107-116).

        Host inputs are streamed: the batch is cut into H2D_CHUNK-image chunks and the copy of chunk
        i+1 (second stream, two staging buffers) overlaps the encoder kernels of chunk i, so PCIe
        time hides behind compute instead of adding to it (pin the tensor for the copy to be async)."""
        a = self.arch
        if pixel_values.dim() != 4 or tuple(pixel_values.shape[1:]) != (3, a.image, a.image):
            raise ValueError(f"pixel_values must be [B,3,{a.image},{a.image}], got {tuple(pixel_values.shape)}")
        b = pixel_values.shape[0]
        out = torch.empty((b, a.proj_dim), dtype=torch.float32, device=self.device)
        if b == 0:
            return out
        if pixel_values.is_cuda or b <= self.H2D_CHUNK:
            pv = pixel_values.to(device=self.device, dtype=torch.float32).contiguous()
            self._encode_images_dev(pv, out, normalize)
            return out
        src = pixel_values.to(dtype=torch.float32).contiguous()
        cs = self.H2D_CHUNK
        if self._stage is None:
            self._stage = [torch.empty((cs, 3, a.image, a.image), dtype=torch.float32, device=self.device)
                           for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._ev_ready = [torch.cuda.Event() for _ in range(2)]
            self._ev_free = [torch.cuda.Event() for _ in range(2)]
        cur = torch.cuda.current_stream(self.device)
        for k in range(2):
            self._ev_free[k].record(cur)  # both staging buffers are free as of now on this stream
        # the first chunk's copy cannot hide behind compute, so it is a small one
        bounds, b0 = [], 0
        while b0 < b:
            n = min(min(self.H2D_FIRST_CHUNK, cs) if b0 == 0 else cs, b - b0)
            bounds.append((b0, n))
            b0 += n
        for i, (b0, n) in enumerate(bounds):
            k = i & 1
            self._copy_stream.wait_event(self._ev_free[k])
            with torch.cuda.stream(self._copy_stream):
                self._stage[k][:n].copy_(src[b0:b0 + n], non_blocking=True)
                self._ev_ready[k].record(self._copy_stream)
            cur.wait_event(self._ev_ready[k])
            self._encode_images_dev(self._stage[k][:n], out[b0:b0 + n], normalize)
            self._ev_free[k].record(cur)
        return out

    BUCKET_MIN_BATCH = 64      # below this a text batch is one pass at its longest caption
    BUCKET_STEP = 16           # bucket tops are multiples of this (the attention kernel's key granularity)
    BUCKET_MIN_TOKENS = 24576  # rows x positions a pass needs to stay out of the launch-bound regime

    def _bucket_tops(self, lens: torch.Tensor, width: int) -> torch.Tensor:
        """Per caption, the number of positions its pass covers.  Captions are grouped by length rounded up
        to BUCKET_STEP; groups are merged upwards until every pass has BUCKET_MIN_TOKENS rows x positions
        (a pass is ~100 launches: measured, five ~200-row passes are slower than one padded pass)."""
        step = self.BUCKET_STEP
        tops = ((lens + step - 1) // step * step).clamp(max=width)
        levels = sorted(set(tops.tolist()))
        counts = {t: int((tops == t).sum()) for t in levels}
        remap, pending, pending_rows = {}, [], 0
        for i, t in enumerate(levels):
            pending.append(t)
            pending_rows += counts[t]
            # close the group at t once it is big enough; smaller groups ride along with the next longer one
            # (the last group is served as it is: a short launch-bound pass costs less than lengthening the rest)
            if i == len(levels) - 1 or pending_rows * t >= self.BUCKET_MIN_TOKENS:
                for p_ in pending:
                    remap[p_] = t
                pending, pending_rows = [], 0
        return torch.tensor([remap[t] for t in tops.tolist()], dtype=torch.int64)

    def encode_texts(self, input_ids: torch.Tensor, normalize: bool = True,
                     lengths: Optional[Union[torch.Tensor, List[int]]] = None, bucket: bool = True) -> torch.Tensor:
        """[B, L<=77] int (right padded or not) -> [B, proj_dim] fp32 on the GPU.

        `lengths` (host ints, tokens per caption including its EOS -- what the tokenizer's attention_mask
        sums to; derived here when `input_ids` is a host tensor) lets the batch run **length-bucketed**:
        rows are grouped by length and every group is encoded on its first `top` positions only.  The text tower is causal and pooled at the first EOS,
        so positions after it never reach the output (encode_text row), and the reference
        itself encodes a single caption unpadded (models/clip_model.py:133-138).  Without `lengths` (device
        ids) or with bucket=False every row runs on all L positions (padded to the context length with EOS
        if L is shorter).  Pays off when captions are short (one pass at 16 positions instead of 77); on
        lengths spread over 3..77 it is a wash against the padded pass (profiles/r1_notes.md)."""
        a = self.arch
        if input_ids.dim() != 2 or input_ids.shape[1] > a.context:
            raise ValueError(f"input_ids must be [B, L<={a.context}], got {tuple(input_ids.shape)}")
        if not bucket:
            lengths = None
        elif lengths is None and not input_ids.is_cuda and input_ids.shape[0] > 0:
            # host ids (the tokenizer's output): the lengths are one cheap host pass away
            is_eos = input_ids == a.eos_id
            first = is_eos.int().argmax(dim=1) + 1
            lengths = torch.where(is_eos.any(dim=1), first, torch.full_like(first, input_ids.shape[1]))
        ids = input_ids.to(device=self.device, dtype=torch.int32, non_blocking=True)
        b, l = ids.shape
        out = torch.empty((b, a.proj_dim), dtype=torch.float32, device=self.device)
        if b == 0:
            return out
        if lengths is not None:
            lens = torch.as_tensor(lengths, dtype=torch.int64, device="cpu").reshape(-1)
            if lens.numel() != b:
                raise ValueError(f"lengths must have one entry per caption ({b}), got {lens.numel()}")
            lens = lens.clamp(1, l)
            if b < self.BUCKET_MIN_BATCH:
                step = self.BUCKET_STEP  # multiples of 16 only: few distinct shapes (and CUDA graphs) per batch size
                top = min(a.context, (int(lens.max()) + step - 1) // step * step)
                sub = ids[:, :min(top, l)]
                if top > l:  # ids narrower than the pass (tokenizer padded to the longest caption): pad with EOS
                    sub = torch.cat([sub, torch.full((b, top - l), a.eos_id, dtype=torch.int32, device=self.device)], dim=1)
                self._run_tower("text", sub.contiguous(), out, normalize)
                return out
            tops = self._bucket_tops(lens, l)
            for top in sorted(set(tops.tolist())):
                rows = torch.nonzero(tops == top).reshape(-1).to(self.device, non_blocking=True)
                sub = ids.index_select(0, rows)[:, :top].contiguous()
                sub_out = torch.empty((rows.numel(), a.proj_dim), dtype=torch.float32, device=self.device)
                self._run_tower("text", sub, sub_out, normalize)
                out.index_copy_(0, rows, sub_out)
            return out
        if l < a.context:
            pad = torch.full((b, a.context - l), a.eos_id, dtype=torch.int32, device=self.device)
            ids = torch.cat([ids, pad], dim=1)
        self._run_tower("text", ids.contiguous(), out, normalize)
        return out

    # transformers-4.x style accessors used by the reference (tensor in, tensor out)
    def get_image_features(self, pixel_values: torch.Tensor, **_) -> torch.Tensor:
        return self.encode_images(pixel_values, normalize=False)

    def get_text_features(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                          **_) -> torch.Tensor:
        lengths = None
        if attention_mask is not None:
            # positions masked out are padding: replace by EOS so the first-EOS pooling row is kept
            input_ids = torch.where(attention_mask.to(input_ids.device).bool(), input_ids,
                                    torch.full_like(input_ids, self.arch.eos_id))
            m = attention_mask
            if not m.is_cuda and bool((m[:, 1:] <= m[:, :-1]).all()):  # a host prefix mask (the tokenizer's)
                lengths = m.sum(dim=1)
        return self.encode_texts(input_ids, normalize=False, lengths=lengths)


# ------------------------------------------------------------------------------------------
# processor (host side; out of the timed path)
# ------------------------------------------------------------------------------------------
class FallbackTokenizer:
    """Deterministic stand-in used ONLY when no CLIP BPE vocabulary is on disk (this image has
    none and no network): words are hashed into the BPE id range.  Shapes, BOS/EOS framing,
    truncation and padding follow CLIPTokenizer; ids do not."""

    model_max_length = 77

    def __init__(self, vocab: int = 49408, bos: int = BOS_ID, eos: int = EOS_ID):
        self.vocab, self.bos, self.eos = vocab, bos, eos

    def _ids(self, text: str) -> List[int]:
        import re
        import zlib

        words = re.findall(r"[a-z0-9]+|[^\sa-z0-9]", text.lower())
        return [1 + zlib.crc32(w.encode("utf-8")) % (self.bos - 1) for w in words]

    def __call__(self, texts, padding=True, truncation=True, max_length=None, return_tensors="pt", **_):
        if isinstance(texts, str):
            texts = [texts]
        max_length = max_length or self.model_max_length
        rows = []
        for t in texts:
            ids = [self.bos] + self._ids(t)
            if truncation:
                ids = ids[: max_length - 1]
            rows.append(ids + [self.eos])
        width = max_length if padding == "max_length" else max(len(r) for r in rows)
        input_ids = torch.full((len(rows), width), self.eos, dtype=torch.long)
        mask = torch.zeros((len(rows), width), dtype=torch.long)
        for i, r in enumerate(rows):
            input_ids[i, : len(r)] = torch.tensor(r, dtype=torch.long)
            mask[i, : len(r)] = 1
        return {"input_ids": input_ids, "attention_mask": mask}


class ClmProcessor:
    """Stand-in for transformers.CLIPProcessor with the calls the reference makes:
    processor(images=..., return_tensors="pt") and processor(text=[...], padding=True,
    truncation=True, return_tensors="pt"); `.tokenizer` as used by embed_text.py:35-41."""

    def __init__(self, name: str):
        from transformers import CLIPImageProcessor

        self.image_processor = CLIPImageProcessor()  # CLIP defaults: 224 bicubic, centre crop, mean/std
        self.tokenizer = None
        try:
            from transformers import CLIPTokenizerFast

            tok = CLIPTokenizerFast.from_pretrained(name, local_files_only=True)
            # transformers 5 can hand back an EMPTY tokenizer when no vocabulary files exist
            if tok.eos_token_id != EOS_ID or tok.bos_token_id != BOS_ID or len(tok) < 49408:
                raise RuntimeError("CLIP BPE vocabulary not available")
            self.tokenizer = tok
        except Exception:
            self.tokenizer = FallbackTokenizer()
            print(f"[clip_model] no CLIP BPE vocabulary on disk for '{name}': using the hashed "
                  f"FallbackTokenizer (shapes only; see DESIGN.md)")

    def preprocess_images_gpu(self, images, device: Union[str, torch.device] = "cuda") -> torch.Tensor:
        """The same pixel_values as processor(images=...) but resized / cropped / normalised on the
        GPU (clm_preprocess_images): the decoded uint8 pixels travel over PCIe in ONE packed copy and
        the host never resamples.  `images`: PIL images or uint8 numpy HWC arrays, any sizes."""
        import numpy as np

        from .. import kernels as K

        ip = self.image_processor
        size = ip.size["shortest_edge"] if "shortest_edge" in ip.size else ip.size["height"]
        arrs = []
        for im in (images if isinstance(images, (list, tuple)) else [images]):
            a = np.asarray(im.convert("RGB")) if hasattr(im, "convert") else np.asarray(im)
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
                raise ValueError(f"images must be RGB uint8 [H, W, 3], got {a.shape} {a.dtype}")
            arrs.append(np.ascontiguousarray(a))
        offs, total = [], 0
        for a in arrs:
            offs.append(total)
            total += (a.size + 255) // 256 * 256
        host = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
        hv = host.numpy()
        for a, o in zip(arrs, offs):
            hv[o:o + a.size] = a.reshape(-1)
        devbuf = host.to(device, non_blocking=True)
        views = [devbuf[o:o + a.size].view(a.shape[0], a.shape[1], 3) for a, o in zip(arrs, offs)]
        return K.preprocess_images(views, out_size=size, mean=tuple(ip.image_mean), std=tuple(ip.image_std))

    def __call__(self, text=None, images=None, return_tensors="pt", padding=True, truncation=True, **kw):
        out = {}
        if text is not None:
            enc = self.tokenizer(text, padding=padding, truncation=truncation, return_tensors="pt",
                                 **{k: v for k, v in kw.items() if k == "max_length"})
            out["input_ids"], out["attention_mask"] = enc["input_ids"], enc["attention_mask"]
        if images is not None:
            out["pixel_values"] = self.image_processor(images=images, return_tensors="pt")["pixel_values"]
        return out


# ------------------------------------------------------------------------------------------
# reference functions
# ------------------------------------------------------------------------------------------
def _load_clip_config(config_path: Union[str, Path]) -> dict:
    path = Path(config_path)
    if not path.exists():
        raise FileNotFoundError(f"CLIP config file not found: {path}")
    with open(path, "r", encoding="utf-8") as f:
        return yaml.safe_load(f)


def _get_device(device_str: Optional[str] = None) -> torch.device:
    """Reference rule (models/clip_model.py:23-28): explicit string wins, else CUDA if present.
    The B200 path additionally refuses CPU: there is no fallback implementation."""
    if device_str is None:
        if torch.cuda.is_available():
            return torch.device("cuda")
        raise RuntimeError("no CUDA device: the B200 path has no CPU fallback")
    dev = torch.device(device_str)
    if dev.type != "cuda":
        raise ValueError(f"device '{device_str}' requested but the B200 path only runs on CUDA; "
                         f"set model.device to 'cuda' (or remove it) in the CLIP config")
    return dev


def _get_dtype(dtype_str: str, device: torch.device) -> torch.dtype:
    """The reference picks fp16 on CUDA else fp32 (:31-34).  Here the arithmetic type is fixed:
    bf16 tensor-core operands, fp32 accumulation / LN statistics / softmax; the config key selects the type of
    the residual stream (load_clip_model) and is otherwise reported."""
    return torch.bfloat16


def random_init_state_dict(arch: ClipArch, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init weights of the named architecture with transformers' own init scheme
    (there are no checkpoints on this box; BASELINE.json asks for exactly this)."""
    from transformers import CLIPModel

    torch.manual_seed(seed)
    hf = CLIPModel(hf_config_for(arch))
    hf.eval()
    return {k: v.detach().clone() for k, v in hf.state_dict().items()}


def _find_local_checkpoint(model_name: str) -> Optional[str]:
    """Path of a local copy of `model_name` (a directory holding config.json, or a cached hub snapshot);
    None when nothing is on disk.  Only "nothing on disk" selects random init in load_clip_model."""
    p = Path(model_name)
    if p.is_dir() and (p / "config.json").exists():
        return str(p)
    try:
        from huggingface_hub import try_to_load_from_cache

        hit = try_to_load_from_cache(model_name, "config.json")
        if isinstance(hit, str) and os.path.exists(hit):
            return str(Path(hit).parent)
    except ImportError:
        pass
    return None


def load_clip_model(
    config_path: Union[str, Path] = "config/clip_config.yaml",
    use_lora: bool = False,
    lora_weights_path: Optional[Union[str, Path]] = None,
) -> Tuple[B200ClipModel, ClmProcessor, torch.device]:
    """Load CLIP (+ optional LoRA adapter) per the YAML config; same contract as the reference,
    including its soft fallbacks: a missing/unset LoRA dir prints a notice and continues without
    LoRA (reference :65-75)."""
    config = _load_clip_config(config_path)
    model_cfg = config.get("model", {}) or {}
    model_name = model_cfg.get("name", "openai/clip-vit-base-patch32")
    device = _get_device(model_cfg.get("device"))
    dtype = _get_dtype(model_cfg.get("dtype", "bfloat16"), device)
    seed = int(model_cfg.get("seed", 0))
    print(f"[clip_model] Loading CLIP model '{model_name}' on device: {device} (dtype={dtype})")

    # Weights: a local checkpoint (directory or cached hub snapshot) is loaded and every failure while
    # doing so is raised, as in the reference (models/clip_model.py:59-63) -- a corrupt or unsupported
    # checkpoint must not turn into a model that serves garbage.  Random-init weights of the named
    # architecture are used ONLY when no checkpoint exists on this machine (there is no network here;
    # BASELINE.json measures exactly that) or when the config asks for them (`model.random_init: true`
    # / CLM_RANDOM_INIT=1).
    want_random = bool(model_cfg.get("random_init", False)) or os.environ.get("CLM_RANDOM_INIT") == "1"
    ckpt = None if want_random else _find_local_checkpoint(model_name)
    if ckpt is None:
        arch = arch_from_name(model_name)
        why = "random_init requested" if want_random else "no local checkpoint"
        print(f"[clip_model] {why} for '{model_name}': random-init weights of that architecture (seed={seed})")
        state_dict = random_init_state_dict(arch, seed)
    else:
        from transformers import CLIPModel

        hf = CLIPModel.from_pretrained(ckpt, local_files_only=True)
        arch = arch_from_hf_config(hf.config, model_name)
        state_dict = hf.state_dict()

    lora = None
    if use_lora:
        paths_cfg = config.get("paths", {}) or {}
        if lora_weights_path is None:
            lora_weights_path = paths_cfg.get("lora_weights_dir")
        if lora_weights_path is None:
            print("[clip_model] use_lora=True but lora_weights_path is not set, continuing without LoRA.")
        else:
            lora_path = Path(lora_weights_path)
            if not lora_path.exists():
                print(f"[clip_model] LoRA weights not found at: {lora_path}, continuing without LoRA.")
            else:
                print(f"[clip_model] Loading LoRA weights from: {lora_path}")
                lora = load_lora_adapter(lora_path)

    # model.dtype of the YAML: the reference runs the WHOLE model in fp16 for "float16" on CUDA and in fp32 otherwise
    # (models/clip_model.py:31-34).  Here the GEMM operands are always bf16; the key selects the residual stream:
    # "float32" keeps it in fp32 (the shipped config), a half type stores it in bf16 (CLM_RESIDUAL_DTYPE overrides).
    stream = os.environ.get("CLM_RESIDUAL_DTYPE") or (
        "float32" if str(model_cfg.get("dtype", "bfloat16")).lower() in ("float32", "fp32", "float") else "bfloat16")
    model = B200ClipModel(arch, state_dict, lora=lora, device=device, residual_dtype=stream)
    print(f"[clip_model] residual stream: {model.residual_dtype}")
    processor = ClmProcessor(model_name)
    model.eval()
    return model, processor, device


def encode_image(image_path: Union[str, Path], model: B200ClipModel, processor, device: torch.device) -> torch.Tensor:
    """One image file -> L2-normalised embedding (d,) on CPU fp32 (reference :89-118)."""
    from PIL import Image

    image_path = Path(image_path)
    if not image_path.exists():
        raise FileNotFoundError(f"Image not found: {image_path}")
    image = Image.open(image_path).convert("RGB")
    inputs = processor(images=image, return_tensors="pt")
    pixel_values = inputs["pixel_values"].to(device)
    with torch.no_grad():
        feats = model.encode_images(pixel_values, normalize=True)  # x / ||x|| fused (clm_l2norm)
    return feats.squeeze(0).to("cpu", torch.float32)


def encode_text(text: str, model: B200ClipModel, processor, device: torch.device) -> torch.Tensor:
    """One caption -> L2-normalised embedding (d,) on CPU fp32 (reference :121-150)."""
    inputs = processor(text=[text], return_tensors="pt", padding=True, truncation=True)
    input_ids = inputs["input_ids"].to(device)
    with torch.no_grad():
        feats = model.encode_texts(input_ids, normalize=True)
    return feats.squeeze(0).to("cpu", torch.float32)
