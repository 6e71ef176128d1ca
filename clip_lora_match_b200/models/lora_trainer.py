"""LoRA training step on the B200 kernels — what `loss.backward(); optimizer.step()` of the reference's
scripts/train_lora.py:170-211 runs, for a B200ClipModel with an unmerged LoRA adapter.

    trainer = LoraTrainer(model, lr=1e-4, weight_decay=0.01, max_grad_norm=1.0, temperature=0.07)
    loss = trainer.step(pixel_values, input_ids, lr=...)       # forward + backward + clip + AdamW
    adapter = trainer.export_adapter()                         # PEFT layout (save_lora_adapter / model.set_lora)

Design:
  * The base model is frozen (models/lora_adapter.py:46-56), so the backward only needs the gradient of the
    activations (to reach the adapters of the lower layers) and of the LoRA factors.
  * Parameters live in the FUSED layouts the GEMMs consume: per (tower, fused GEMM) a master A^T [L, in, cols] and
    a master B_cat [L, out_total, cols] in fp32 (cols = total rank of the Linears sharing that GEMM, padded to 64).
    Padding columns and the off-diagonal blocks of B_cat are structural zeros: `grad_mult` is 0 there and the
    optimizer never touches them; elsewhere it carries the factor that turns the GEMM's raw product into the
    gradient of the PEFT parameter (1 for A, alpha/r for B).  All masters, gradients and AdamW moments are slices
    of four flat buffers, so clipping + AdamW is two launches (clm_adamw_step).
  * Every contraction is clm_gemm_epi (tcgen05): forward as in the encoder but layer by layer with the activations
    kept; dgrad = dy W (+ u A as K extension) on pre-transposed copies of the frozen weights; u = dy (sB);
    dB_cat = dy^T t and dA^T = x^T u on transposed activations (clm_transpose_to_bf16), accumulated in fp32 in place.
  * LayerNorm / QuickGELU / attention backward, the InfoNCE loss and the optimizer are the kernels of
    csrc/clm_train.cu.  The whole step is stream ordered with static buffers: after one eager pass it is captured
    into a CUDA graph and replayed (the reference's batch of 8 is ~700 launches of a few microseconds each).
There is no autograd and no CPU fallback.  LoRA dropout is not applied (the step is deterministic); see DESIGN.md.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from .. import _lib
from .._lib import EPI_NONE, EPI_SPLIT_K, OUT_BF16, OUT_F32, check, cur_stream, ptr
from .clip_model import LORA_COLS, B200ClipModel
from .lora_adapter import LoraAdapter

_GROUPS = (
    # name, prefix below the layer, input is, members (leaf, out-features key)
    ("qkv", "self_attn", "width", (("q_proj", "width"), ("k_proj", "width"), ("v_proj", "width"))),
    ("out", "self_attn", "width", (("out_proj", "width"),)),
    ("fc1", "mlp", "width", (("fc1", "mlp"),)),
    ("fc2", "mlp", "mlp", (("fc2", "width"),)),
)


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def allreduce_gradients(grad: torch.Tensor, loss: Optional[torch.Tensor] = None, group=None) -> None:
    """The exchange step of data-parallel training: ONE all-reduce (sum) of the flat gradient buffer -- every rank has
    already scaled its loss by 1 / (grad_accum_steps * world), so the sum IS the mean gradient over the global batch --
    and one 4-byte all-reduce of the scaled loss, which then reads as the mean loss over the ranks.  NCCL on the GPUs
    (NVLink; 4 MB for ViT-B/32 with r = 8), gloo in the CPU tests.  No-op without an initialised process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    if loss is not None:
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)


class _Group:
    """One fused GEMM's adapters over all layers of a tower."""

    def __init__(self, name, in_dim, outs, present, r, cols, layers):
        self.name, self.in_dim, self.outs, self.present = name, in_dim, outs, present
        self.r, self.cols, self.layers = r, cols, layers
        self.out_total = sum(outs)
        self.na = layers * in_dim * cols          # elements of the A^T masters
        self.nb = layers * self.out_total * cols  # elements of the B_cat masters
        self.off_a = self.off_b = 0               # offsets into the flat buffers


class _Tower:
    pass


class LoraTrainer:
    SMALL_WGRAD_ROWS = 2048  # token rows up to which the LoRA weight gradients take the one-launch CUDA-core kernel

    def __init__(self, model: B200ClipModel, lr: float = 1e-4, weight_decay: float = 0.01,
                 max_grad_norm: float = 1.0, temperature: float = 0.07, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, grad_accum_steps: int = 1, use_graph: bool = True, deterministic: bool = False,
                 distributed: bool = False):
        if model.lora is None:
            raise ValueError("LoraTrainer needs a model with a LoRA adapter (attach_lora_to_clip first)")
        self.model = model
        self.arch = model.arch
        self.device = model.device
        self.lib = _lib.load()
        self.lr, self.weight_decay, self.max_grad_norm = float(lr), float(weight_decay), float(max_grad_norm)
        self.temperature, self.betas, self.eps = float(temperature), betas, float(eps)
        self.grad_accum_steps = int(grad_accum_steps)
        self.use_graph = use_graph
        # data-parallel training (an extension: the reference's script is one process): every rank runs the step on
        # its own micro-batch and the flat gradient buffer is all-reduced once before the optimizer; the loss is the
        # mean over ranks of the local InfoNCE losses (negatives are NOT shared across ranks), i.e. exactly what
        # gradient_accumulation_steps = world would compute in one process
        self.world = 1
        if distributed:
            import torch.distributed as dist

            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("distributed=True needs an initialised torch.distributed process group")
            self.world = dist.get_world_size()
        # False: the weight-gradient GEMMs may split K (partial products added through the L2 in arrival order:
        # reproducible to fp32 rounding); True: one work unit per tile, every step reproducible bit for bit
        self.deterministic = deterministic
        self.opt_step = 0      # optimizer steps taken
        self._micro = 0        # micro-batches since the last optimizer step
        self.config = model.lora.config
        self.scaling = model.lora.scaling
        self._towers: Dict[str, _Tower] = {}
        self._build_params(model.lora)
        for kind in ("vision", "text"):
            self._towers[kind] = self._build_tower(kind)
        self._batch = None
        self._graph = None
        self._graph_state = 0  # 0: nothing run yet at this batch, 1: eager pass done, 2: captured
        self.refresh_operands()

    # ---- parameters --------------------------------------------------------------------------------------------
    def _build_params(self, adapter: LoraAdapter) -> None:
        a = self.arch
        self._groups: Dict[Tuple[str, str], _Group] = {}
        r = adapter.config.r
        off = 0
        for kind, ta, pre in (("vision", a.vision, "vision_model"), ("text", a.text, "text_model")):
            dims = {"width": ta.width, "mlp": ta.mlp}
            for name, sub, in_key, members in _GROUPS:
                present = [i for i, (leaf, _) in enumerate(members)
                           if f"{pre}.encoder.layers.0.{sub}.{leaf}" in adapter.weights]
                if not present:
                    continue
                cols = (r * len(present) + LORA_COLS - 1) // LORA_COLS * LORA_COLS
                g = _Group(name, dims[in_key], [dims[k] for _, k in members], present, r, cols, ta.layers)
                g.off_a, off = off, off + g.na
                g.off_b, off = off, off + g.nb
                g.sub, g.pre, g.members = sub, pre, members
                self._groups[(kind, name)] = g
        n = off
        dev = self.device
        self.theta = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.mult = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.hyper = torch.zeros(4, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1185, dtype=torch.float32, device=dev)  # [0]: squared gradient norm; rest: partials
        self.loss_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        # fill the masters and the gradient multipliers from the PEFT-layout adapter (host side, once)
        theta_h = torch.zeros(n, dtype=torch.float32)
        mult_h = torch.zeros(n, dtype=torch.float32)
        for (kind, name), g in self._groups.items():
            at = theta_h[g.off_a:g.off_a + g.na].view(g.layers, g.in_dim, g.cols)
            bt = theta_h[g.off_b:g.off_b + g.nb].view(g.layers, g.out_total, g.cols)
            am = mult_h[g.off_a:g.off_a + g.na].view(g.layers, g.in_dim, g.cols)
            bm = mult_h[g.off_b:g.off_b + g.nb].view(g.layers, g.out_total, g.cols)
            for layer in range(g.layers):
                for slot, i in enumerate(g.present):
                    leaf = g.members[i][0]
                    path = f"{g.pre}.encoder.layers.{layer}.{g.sub}.{leaf}"
                    if path not in adapter.weights:
                        raise ValueError(f"LoRA targets differ between layers: {path} has no adapter")
                    wa, wb = adapter.weights[path]
                    row0 = sum(g.outs[:i])
                    if tuple(wa.shape) != (r, g.in_dim) or tuple(wb.shape) != (g.outs[i], r):
                        raise ValueError(f"LoRA shape mismatch on {path}")
                    at[layer, :, slot * r:(slot + 1) * r] = wa.t()
                    bt[layer, row0:row0 + g.outs[i], slot * r:(slot + 1) * r] = wb
                    am[layer, :, slot * r:(slot + 1) * r] = 1.0
                    bm[layer, row0:row0 + g.outs[i], slot * r:(slot + 1) * r] = self.scaling
        self.theta.copy_(theta_h)
        self.mult.copy_(mult_h)
        # bf16 operands of the GEMMs, refreshed from the masters after every optimizer step
        self._ops: Dict[Tuple[str, str], Dict[str, torch.Tensor]] = {}
        for key, g in self._groups.items():
            e = lambda *s: torch.zeros(s, dtype=torch.bfloat16, device=dev)  # noqa: E731
            self._ops[key] = {"a_cat": e(g.layers, g.cols, g.in_dim), "a_catT": e(g.layers, g.in_dim, g.cols),
                              "b_cat": e(g.layers, g.out_total, g.cols), "b_catT": e(g.layers, g.cols, g.out_total)}

    def num_trainable_parameters(self) -> int:
        return int((self.mult != 0).sum().item())

    def _master(self, key, which: str, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        g = self._groups[key]
        buf = self.theta if buf is None else buf
        if which == "a":
            return buf[g.off_a:g.off_a + g.na].view(g.layers, g.in_dim, g.cols)
        return buf[g.off_b:g.off_b + g.nb].view(g.layers, g.out_total, g.cols)

    def refresh_operands(self) -> None:
        """bf16 GEMM operands from the fp32 masters: A_cat, A_cat^T, s B_cat, (s B_cat)^T for every group."""
        lib, st = self.lib, cur_stream()
        for key, g in self._groups.items():
            ops = self._ops[key]
            a, b = self._master(key, "a"), self._master(key, "b")
            check(lib.clm_cast_to_bf16(ptr(a), ptr(ops["a_catT"]), g.na, 1.0, st), "clm_cast_to_bf16")
            check(lib.clm_transpose_to_bf16(ptr(a), 1, g.cols, g.in_dim * g.cols, g.in_dim, g.cols, ptr(ops["a_cat"]),
                                            g.in_dim, g.cols * g.in_dim, g.layers, 1.0, st), "clm_transpose_to_bf16")
            check(lib.clm_cast_to_bf16(ptr(b), ptr(ops["b_cat"]), g.nb, self.scaling, st), "clm_cast_to_bf16")
            check(lib.clm_transpose_to_bf16(ptr(b), 1, g.cols, g.out_total * g.cols, g.out_total, g.cols,
                                            ptr(ops["b_catT"]), g.out_total, g.cols * g.out_total, g.layers,
                                            self.scaling, st), "clm_transpose_to_bf16")

    def export_adapter(self) -> LoraAdapter:
        """The current LoRA factors in PEFT layout (path -> (A [r, in], B [out, r]) fp32 on the CPU)."""
        weights = {}
        for key, g in self._groups.items():
            at, bt = self._master(key, "a").cpu(), self._master(key, "b").cpu()
            for layer in range(g.layers):
                for slot, i in enumerate(g.present):
                    leaf = g.members[i][0]
                    row0 = sum(g.outs[:i])
                    weights[f"{g.pre}.encoder.layers.{layer}.{g.sub}.{leaf}"] = (
                        at[layer, :, slot * g.r:(slot + 1) * g.r].t().contiguous(),
                        bt[layer, row0:row0 + g.outs[i], slot * g.r:(slot + 1) * g.r].contiguous())
        return LoraAdapter(config=self.config, weights=weights,
                           base_model_name_or_path=self.model.lora.base_model_name_or_path)

    def gradients(self) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
        """Accumulated gradients in PEFT layout (path -> (dA, dB)) -- what `p.grad` holds in the reference after
        loss.backward().  For tests and inspection."""
        out = {}
        gm = self.grad * self.mult
        for key, g in self._groups.items():
            at, bt = self._master(key, "a", gm).cpu(), self._master(key, "b", gm).cpu()
            for layer in range(g.layers):
                for slot, i in enumerate(g.present):
                    leaf = g.members[i][0]
                    row0 = sum(g.outs[:i])
                    out[f"{g.pre}.encoder.layers.{layer}.{g.sub}.{leaf}"] = (
                        at[layer, :, slot * g.r:(slot + 1) * g.r].t().contiguous(),
                        bt[layer, row0:row0 + g.outs[i], slot * g.r:(slot + 1) * g.r].contiguous())
        return out

    def sync_model(self) -> None:
        """Hand the trained factors to the inference model (rebuilds its towers)."""
        self.model.set_lora(self.export_adapter())

    # ---- frozen weights ----------------------------------------------------------------------------------------
    def _build_tower(self, kind: str) -> _Tower:
        a, sd, dev = self.arch, self.model._sd, self.device
        ta = a.vision if kind == "vision" else a.text
        pre = "vision_model" if kind == "vision" else "text_model"
        t = _Tower()
        t.kind, t.ta, t.pre = kind, ta, pre
        t.tokens = a.vision_tokens if kind == "vision" else a.context
        f32 = lambda x: x.to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        bf = lambda x: x.to(device=dev, dtype=torch.bfloat16).contiguous()  # noqa: E731
        t.layers = []
        for i in range(ta.layers):
            lp = f"{pre}.encoder.layers.{i}"
            ap = f"{lp}.self_attn"
            wqkv = torch.cat([sd[f"{ap}.q_proj.weight"], sd[f"{ap}.k_proj.weight"], sd[f"{ap}.v_proj.weight"]], 0)
            L = {
                "ln1_g": f32(sd[f"{lp}.layer_norm1.weight"]), "ln1_b": f32(sd[f"{lp}.layer_norm1.bias"]),
                "ln2_g": f32(sd[f"{lp}.layer_norm2.weight"]), "ln2_b": f32(sd[f"{lp}.layer_norm2.bias"]),
                "w_qkv": bf(wqkv), "w_qkvT": bf(wqkv.t()),
                "b_qkv": f32(torch.cat([sd[f"{ap}.q_proj.bias"], sd[f"{ap}.k_proj.bias"], sd[f"{ap}.v_proj.bias"]], 0)),
                "w_out": bf(sd[f"{ap}.out_proj.weight"]), "w_outT": bf(sd[f"{ap}.out_proj.weight"].t()),
                "b_out": f32(sd[f"{ap}.out_proj.bias"]),
                "w_fc1": bf(sd[f"{lp}.mlp.fc1.weight"]), "w_fc1T": bf(sd[f"{lp}.mlp.fc1.weight"].t()),
                "b_fc1": f32(sd[f"{lp}.mlp.fc1.bias"]),
                "w_fc2": bf(sd[f"{lp}.mlp.fc2.weight"]), "w_fc2T": bf(sd[f"{lp}.mlp.fc2.weight"].t()),
                "b_fc2": f32(sd[f"{lp}.mlp.fc2.bias"]),
            }
            t.layers.append(L)
        t.pos_emb = f32(sd[f"{pre}.embeddings.position_embedding.weight"])
        if kind == "vision":
            k = 3 * a.patch * a.patch
            t.kpad = (k + 63) // 64 * 64
            t.np = (a.image // a.patch) ** 2
            pw = torch.zeros((ta.width, t.kpad), dtype=torch.float32)
            pw[:, :k] = sd[f"{pre}.embeddings.patch_embedding.weight"].reshape(ta.width, k)
            t.patch_w = bf(pw)
            t.class_emb = f32(sd[f"{pre}.embeddings.class_embedding"])
            t.pre_ln_g, t.pre_ln_b = f32(sd[f"{pre}.pre_layrnorm.weight"]), f32(sd[f"{pre}.pre_layrnorm.bias"])
            t.final_g, t.final_b = f32(sd[f"{pre}.post_layernorm.weight"]), f32(sd[f"{pre}.post_layernorm.bias"])
            proj = sd["visual_projection.weight"]
        else:
            t.tok_emb = f32(sd[f"{pre}.embeddings.token_embedding.weight"])
            t.final_g, t.final_b = f32(sd[f"{pre}.final_layer_norm.weight"]), f32(sd[f"{pre}.final_layer_norm.bias"])
            proj = sd["text_projection.weight"]
        t.proj_w, t.proj_wT = bf(proj), bf(proj.t())
        t.groups = {name: self._groups[(kind, name)] for name, *_ in _GROUPS if (kind, name) in self._groups}
        return t

    # ---- activations / scratch for a batch size ----------------------------------------------------------------
    def _alloc(self, batch: int) -> None:
        if self._batch == batch:
            return
        a, dev = self.arch, self.device
        self._batch, self._graph, self._graph_state = batch, None, 0
        bfz = lambda *s: torch.zeros(s, dtype=torch.bfloat16, device=dev)  # noqa: E731
        f32z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731
        self.pixel_values = f32z(batch, 3, a.image, a.image)
        self.input_ids = torch.full((batch, a.context), a.eos_id, dtype=torch.int32, device=dev)
        for t in self._towers.values():
            ta, D, M = t.ta, t.ta.width, t.ta.mlp
            rows = batch * t.tokens
            t.rows, t.rows_pad = rows, _pad8(rows)
            Ln = ta.layers
            t.H = f32z(2 * Ln + 1, rows, D)           # residual stream before / inside / after every layer
            t.x1, t.x2, t.ao = bfz(Ln, rows, D), bfz(Ln, rows, D), bfz(Ln, rows, D)
            t.qkv = bfz(Ln, rows, 3 * D)
            t.z, t.g = bfz(Ln, rows, M), bfz(Ln, rows, M)
            t.t = {n: bfz(Ln, rows, g.cols) for n, g in t.groups.items()}
            cmax = max([g.cols for g in t.groups.values()] + [LORA_COLS])
            wide = max(3 * D, M)
            t.dh, t.dh_bf = f32z(rows, D), bfz(rows, D)
            t.dg, t.dqkv = bfz(rows, M), bfz(rows, 3 * D)  # dg (dz in place), dq | dk | dv
            t.dx, t.dao = bfz(rows, D), bfz(rows, D)
            t.u = bfz(rows, cmax)
            t.dyT, t.xT = bfz(wide, t.rows_pad), bfz(wide, t.rows_pad)
            t.tT, t.uT = bfz(cmax, t.rows_pad), bfz(cmax, t.rows_pad)
            nb = self.lib.clm_attention_bwd_scratch_bytes(batch, t.tokens, ta.heads)
            t.att_scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
            t.pooled, t.dpooled = bfz(batch, D), bfz(batch, D)
            t.feat, t.dfeat = f32z(batch, a.proj_dim), f32z(batch, a.proj_dim)
            t.dfeat_bf = bfz(batch, a.proj_dim)
            t.eos = torch.zeros(batch, dtype=torch.int32, device=dev)
            if t.kind == "vision":
                t.patches = bfz(batch * t.np, t.kpad)
                t.patch_out = f32z(batch * t.np, D)
        nb = self.lib.clm_clip_loss_workspace_bytes(batch, a.proj_dim)
        self.loss_ws = torch.empty(nb, dtype=torch.uint8, device=dev)

    # ---- thin launch helpers -----------------------------------------------------------------------------------
    def _gemm(self, a, w, out, bias=None, a2=None, w2=None, accumulate=False, M=None, N=None, K=None, K2=None,
              split_k=False):
        M = a.shape[0] if M is None else M
        K = a.shape[1] if K is None else K
        N = w.shape[0] if N is None else N
        k2 = 0
        if a2 is not None:
            k2 = a2.shape[1] if K2 is None else K2
        od = OUT_F32 if out.dtype == torch.float32 else OUT_BF16
        res = out if accumulate else None
        check(self.lib.clm_gemm_epi(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K,
                                    ptr(a2), a2.stride(0) if a2 is not None else 0,
                                    ptr(w2), w2.stride(0) if w2 is not None else 0, k2,
                                    ptr(out), out.stride(0), od, ptr(bias), ptr(res),
                                    out.stride(0) if accumulate else 0,
                                    EPI_SPLIT_K if (split_k and not self.deterministic) else EPI_NONE, cur_stream()),
              "clm_gemm_epi")

    def _transpose(self, src, dst, rows, cols):
        check(self.lib.clm_transpose_to_bf16(ptr(src), 0, src.stride(0), 0, rows, cols, ptr(dst), dst.stride(0), 0, 1,
                                             1.0, cur_stream()), "clm_transpose_to_bf16")

    def _ln(self, x, g, b, y, rows, D):
        check(self.lib.clm_layernorm(ptr(x), ptr(g), ptr(b), ptr(y), rows, D, self.arch.ln_eps, cur_stream()),
              "clm_layernorm")

    def _ln_bwd(self, dy, x, g, t, rows, D, accumulate=True):
        check(self.lib.clm_layernorm_bwd(ptr(dy), 0, ptr(x), ptr(g), ptr(t.dh), ptr(t.dh_bf), rows, D,
                                         self.arch.ln_eps, int(accumulate), None, 0, 0, cur_stream()),
              "clm_layernorm_bwd")

    # ---- forward -----------------------------------------------------------------------------------------------
    def _linear_fwd(self, t: _Tower, layer: int, name: str, x, w, bias, out, accumulate=False):
        """out (=|+=) x W^T + (x A_cat^T)(s B_cat)^T + bias, keeping t = x A_cat^T for the backward."""
        g = t.groups.get(name)
        if g is None:
            self._gemm(x, w, out, bias=bias, accumulate=accumulate)
            return
        ops = self._ops[(t.kind, name)]
        tt = t.t[name][layer]
        self._gemm(x, ops["a_cat"][layer], tt)
        self._gemm(x, w, out, bias=bias, a2=tt, w2=ops["b_cat"][layer], accumulate=accumulate)

    def _forward_tower(self, t: _Tower, batch: int) -> None:
        a, lib, st = self.arch, self.lib, cur_stream()
        ta, D, rows = t.ta, t.ta.width, t.rows
        if t.kind == "vision":
            check(lib.clm_patch_im2col(ptr(self.pixel_values), ptr(t.patches), batch, a.image, a.patch, t.kpad, st),
                  "clm_patch_im2col")
            self._gemm(t.patches, t.patch_w, t.patch_out)
            check(lib.clm_vision_embed_ln(ptr(t.patch_out), ptr(t.class_emb), ptr(t.pos_emb), ptr(t.pre_ln_g),
                                          ptr(t.pre_ln_b), ptr(t.H[0]), batch, t.np, D, a.ln_eps, st),
                  "clm_vision_embed_ln")
        else:
            check(lib.clm_embed_text(ptr(self.input_ids), ptr(t.tok_emb), ptr(t.pos_emb), ptr(t.H[0]), ptr(t.eos),
                                     batch, t.tokens, D, a.vocab, a.eos_id, st), "clm_embed_text")
        for l, L in enumerate(t.layers):
            h_in, h_mid, h_out = t.H[2 * l], t.H[2 * l + 1], t.H[2 * l + 2]
            self._ln(h_in, L["ln1_g"], L["ln1_b"], t.x1[l], rows, D)
            self._linear_fwd(t, l, "qkv", t.x1[l], L["w_qkv"], L["b_qkv"], t.qkv[l])
            check(lib.clm_attention(ptr(t.qkv[l]), ptr(t.ao[l]), batch, t.tokens, ta.heads, int(t.kind == "text"), st),
                  "clm_attention")
            h_mid.copy_(h_in)
            self._linear_fwd(t, l, "out", t.ao[l], L["w_out"], L["b_out"], h_mid, accumulate=True)
            self._ln(h_mid, L["ln2_g"], L["ln2_b"], t.x2[l], rows, D)
            self._linear_fwd(t, l, "fc1", t.x2[l], L["w_fc1"], L["b_fc1"], t.z[l])
            check(lib.clm_quickgelu_fwd(ptr(t.z[l]), ptr(t.g[l]), rows * ta.mlp, st), "clm_quickgelu_fwd")
            h_out.copy_(h_mid)
            self._linear_fwd(t, l, "fc2", t.g[l], L["w_fc2"], L["b_fc2"], h_out, accumulate=True)
        idx = ptr(t.eos) if t.kind == "text" else None
        check(lib.clm_pool_ln(ptr(t.H[2 * ta.layers]), idx, ptr(t.final_g), ptr(t.final_b), ptr(t.pooled), batch,
                              t.tokens, D, a.ln_eps, st), "clm_pool_ln")
        self._gemm(t.pooled, t.proj_w, t.feat)

    # ---- backward ----------------------------------------------------------------------------------------------
    def _linear_bwd(self, t: _Tower, layer: int, name: str, dy, n_out: int, x, in_dim: int, wT, dx, need_dx=True):
        """dx = dy W + (dy sB_cat) A_cat; accumulates dB_cat += dy^T t and dA^T += x^T u into the flat gradient."""
        g = t.groups.get(name)
        rows = t.rows
        if g is None:
            if need_dx:
                self._gemm(dy, wT, dx, M=rows, N=in_dim, K=n_out)
            return
        key = (t.kind, name)
        ops = self._ops[key]
        u = t.u[:, :g.cols]
        self._gemm(dy, ops["b_catT"][layer], u, M=rows, N=g.cols, K=n_out)
        if need_dx:
            self._gemm(dy, wT, dx, a2=u, w2=ops["a_catT"][layer], M=rows, N=in_dim, K=n_out, K2=g.cols)
        gb = self._master(key, "b", self.grad)[layer]
        ga = self._master(key, "a", self.grad)[layer]
        if rows <= self.SMALL_WGRAD_ROWS:
            # few token rows: one CUDA-core launch instead of four transposes + two deep-K GEMMs (launch bound there)
            tt = t.t[name][layer]
            check(self.lib.clm_lora_wgrad_small(ptr(dy), dy.stride(0), n_out, ptr(tt), tt.stride(0), ptr(x), x.stride(0),
                                                in_dim, ptr(u), u.stride(0), g.cols, rows, ptr(gb), ptr(ga),
                                                int(self.deterministic), cur_stream()), "clm_lora_wgrad_small")
            return
        self._transpose(dy, t.dyT, rows, n_out)
        self._transpose(t.t[name][layer], t.tT, rows, g.cols)
        self._gemm(t.dyT, t.tT, gb, accumulate=True, M=n_out, N=g.cols, K=rows, split_k=True)
        self._transpose(x, t.xT, rows, in_dim)
        self._transpose(u, t.uT, rows, g.cols)
        self._gemm(t.xT, t.uT, ga, accumulate=True, M=in_dim, N=g.cols, K=rows, split_k=True)

    def _backward_tower(self, t: _Tower, batch: int) -> None:
        a, lib, st = self.arch, self.lib, cur_stream()
        ta, D, M, rows = t.ta, t.ta.width, t.ta.mlp, t.rows
        causal = int(t.kind == "text")
        # head: dfeat -> projection -> final LayerNorm of the pooled row, scattered into a zeroed dh
        self._gemm(t.dfeat_bf, t.proj_wT, t.dpooled)
        t.dh.zero_()
        t.dh_bf.zero_()
        idx = ptr(t.eos) if t.kind == "text" else None
        check(lib.clm_layernorm_bwd(ptr(t.dpooled), 0, ptr(t.H[2 * ta.layers]), ptr(t.final_g), ptr(t.dh),
                                    ptr(t.dh_bf), batch, D, a.ln_eps, 0, idx, t.tokens, 1, st), "clm_layernorm_bwd")
        for l in range(ta.layers - 1, -1, -1):
            L = t.layers[l]
            h_in, h_mid = t.H[2 * l], t.H[2 * l + 1]
            # h_out = h_mid + fc2(quickgelu(fc1(LN2(h_mid))))
            self._linear_bwd(t, l, "fc2", t.dh_bf, D, t.g[l], M, L["w_fc2T"], t.dg)
            check(lib.clm_quickgelu_bwd(ptr(t.dg), ptr(t.z[l]), ptr(t.dg), rows * M, st), "clm_quickgelu_bwd")
            self._linear_bwd(t, l, "fc1", t.dg, M, t.x2[l], D, L["w_fc1T"], t.dx)
            self._ln_bwd(t.dx, h_mid, L["ln2_g"], t, rows, D)
            # h_mid = h_in + out_proj(attention(qkv(LN1(h_in))))
            self._linear_bwd(t, l, "out", t.dh_bf, D, t.ao[l], D, L["w_outT"], t.dao)
            check(lib.clm_attention_bwd(ptr(t.qkv[l]), ptr(t.dao), ptr(t.dqkv), ptr(t.att_scratch),
                                        t.att_scratch.numel(), batch, t.tokens, ta.heads, causal, st),
                  "clm_attention_bwd")
            need_dx = l > 0  # nothing trainable sits below the first layer's LayerNorm
            self._linear_bwd(t, l, "qkv", t.dqkv, 3 * D, t.x1[l], D, L["w_qkvT"], t.dx, need_dx=need_dx)
            if need_dx:
                self._ln_bwd(t.dx, h_in, L["ln1_g"], t, rows, D)

    # ---- the step ----------------------------------------------------------------------------------------------
    def _loss(self, batch: int, with_grad: bool) -> None:
        v, x = self._towers["vision"], self._towers["text"]
        check(self.lib.clm_clip_loss(ptr(v.feat), ptr(x.feat), batch, self.arch.proj_dim, self.temperature,
                                     1.0 / (self.grad_accum_steps * self.world) if with_grad else 1.0, ptr(self.loss_dev),
                                     ptr(v.dfeat) if with_grad else None, ptr(x.dfeat) if with_grad else None,
                                     ptr(v.dfeat_bf) if with_grad else None, ptr(x.dfeat_bf) if with_grad else None,
                                     ptr(self.loss_ws), self.loss_ws.numel(), cur_stream()), "clm_clip_loss")

    def _forward_backward(self, batch: int) -> None:
        for t in self._towers.values():
            self._forward_tower(t, batch)
        self._loss(batch, True)
        for t in self._towers.values():
            self._backward_tower(t, batch)

    def _optimizer(self) -> None:
        check(self.lib.clm_adamw_step(ptr(self.theta), ptr(self.grad), ptr(self.mult), ptr(self.exp_avg),
                                      ptr(self.exp_avg_sq), self.theta.numel(), ptr(self.hyper), ptr(self.sumsq),
                                      self.max_grad_norm, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                      cur_stream()), "clm_adamw_step")
        self.refresh_operands()
        self.grad.zero_()

    def _load_inputs(self, pixel_values: torch.Tensor, input_ids: torch.Tensor,
                     attention_mask: Optional[torch.Tensor]) -> int:
        a = self.arch
        b = pixel_values.shape[0]
        if tuple(pixel_values.shape[1:]) != (3, a.image, a.image):
            raise ValueError(f"pixel_values must be [B,3,{a.image},{a.image}], got {tuple(pixel_values.shape)}")
        if input_ids.dim() != 2 or input_ids.shape[0] != b or input_ids.shape[1] > a.context:
            raise ValueError(f"input_ids must be [{b}, L<={a.context}], got {tuple(input_ids.shape)}")
        if b < 1:
            raise ValueError("empty batch")
        self._alloc(b)
        self.pixel_values.copy_(pixel_values.to(dtype=torch.float32), non_blocking=True)
        ids = input_ids
        if attention_mask is not None:  # masked-out positions are padding: EOS keeps the first-EOS pooling row
            ids = torch.where(attention_mask.to(ids.device).bool(), ids, torch.full_like(ids, a.eos_id))
        self.input_ids.fill_(a.eos_id)
        self.input_ids[:, :ids.shape[1]].copy_(ids.to(dtype=torch.int32), non_blocking=True)
        return b

    def _set_hyper(self, lr: float) -> None:
        # a pageable source: the copy is staged before the call returns, so the next step may overwrite the values
        t = self.opt_step + 1
        self.hyper.copy_(torch.tensor([lr, 1.0 / (1.0 - self.betas[0] ** t),
                                       1.0 / math.sqrt(1.0 - self.betas[1] ** t), 0.0], dtype=torch.float32))

    def forward_backward(self, pixel_values: torch.Tensor, input_ids: torch.Tensor,
                         attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One micro-batch: loss / grad_accum_steps and its gradient accumulated into the flat buffer
        (train_lora.py:172-187).  Returns the device scalar holding the (scaled) loss."""
        b = self._load_inputs(pixel_values, input_ids, attention_mask)
        self._forward_backward(b)
        self._micro += 1
        return self.loss_dev

    def optimizer_step(self, lr: Optional[float] = None) -> None:
        """clip_grad_norm_ + AdamW + zero_grad (train_lora.py:190-193) and the refresh of the bf16 operands."""
        self._set_hyper(self.lr if lr is None else float(lr))
        if self.world > 1:
            allreduce_gradients(self.grad)
        self._optimizer()
        self.opt_step += 1
        self._micro = 0

    def step(self, pixel_values: torch.Tensor, input_ids: torch.Tensor,
             attention_mask: Optional[torch.Tensor] = None, lr: Optional[float] = None) -> torch.Tensor:
        """forward + backward + optimizer step for grad_accum_steps == 1, replayed from a CUDA graph once the batch
        size has been seen twice.  Returns the device scalar with the loss of this batch (read it with .item())."""
        if self.grad_accum_steps != 1:
            raise ValueError("step() is the fused path for grad_accum_steps == 1; use forward_backward / optimizer_step")
        b = self._load_inputs(pixel_values, input_ids, attention_mask)
        self._set_hyper(self.lr if lr is None else float(lr))
        def exchange():
            if self.world > 1:
                allreduce_gradients(self.grad, self.loss_dev)

        if not self.use_graph or self.lib.clm_prof_is_enabled():
            self._forward_backward(b)
            exchange()
            self._optimizer()
        elif self._graph_state == 0:
            self._forward_backward(b)  # eager once: one-time cudaFuncSetAttribute calls must not fall into a capture
            exchange()
            self._optimizer()
            self._graph_state = 1
        else:
            if self._graph_state == 1:
                # two graphs -- forward + backward, then clip + AdamW + operand refresh -- with the gradient all-reduce
                # (data-parallel training) between them on the same stream
                n0 = self.lib.clm_launch_count()
                g_fb, g_opt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_fb, capture_error_mode="thread_local"):
                    self._forward_backward(b)
                with torch.cuda.graph(g_opt, capture_error_mode="thread_local"):
                    self._optimizer()
                self._graph_launches = self.lib.clm_launch_count() - n0
                self.lib.clm_launch_count_add(-self._graph_launches)
                self._graph, self._graph_opt, self._graph_state = g_fb, g_opt, 2
            self._graph.replay()
            exchange()
            self._graph_opt.replay()
            self.lib.clm_launch_count_add(self._graph_launches)
        self.opt_step += 1
        return self.loss_dev

    @torch.no_grad()
    def eval_loss(self, pixel_values: torch.Tensor, input_ids: torch.Tensor,
                  attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Validation loss of a batch with the current factors (train_lora.py:213-231): forward only."""
        b = self._load_inputs(pixel_values, input_ids, attention_mask)
        for t in self._towers.values():
            self._forward_tower(t, b)
        self._loss(b, False)
        return self.loss_dev

    def features(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Un-normalised image / text features of the last forward (get_image_features / get_text_features)."""
        return self._towers["vision"].feat, self._towers["text"].feat
