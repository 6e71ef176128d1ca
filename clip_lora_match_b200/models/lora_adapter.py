"""LoRA injection and export — the B200 mirror of the reference's models/lora_adapter.py.

Reference surface kept (same names, argument meaning, defaults):
  create_lora_config(config_path)           reference models/lora_adapter.py:21-43
  attach_lora_to_clip(model, lora_config)   reference models/lora_adapter.py:46-56
plus what the reference gets from `peft` and this repo must provide itself (peft is not a
dependency here): the LoraConfig value type, PEFT's target-module matching rule, PEFT's
adapter init (A kaiming-uniform, B zero) and the PEFT on-disk layout
(`adapter_config.json` + `adapter_model.safetensors`, written by `model.save_pretrained`
at reference scripts/train_lora.py:243-247 and read by `PeftModel.from_pretrained` at
reference models/clip_model.py:78).

Nothing here computes on the hot path: adapters are plain CPU fp32 tensors that
clip_model.B200ClipModel packs into the fused-QKV / out_proj K-extension operands.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Tuple, Union

import torch
import yaml

SUPPORTED_TARGETS = ("q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2")
ADAPTER_CONFIG = "adapter_config.json"
ADAPTER_WEIGHTS = "adapter_model.safetensors"


def _load_lora_config(config_path: Union[str, Path]) -> dict:
    path = Path(config_path)
    if not path.exists():
        raise FileNotFoundError(f"LoRA config file not found: {path}")
    with open(path, "r", encoding="utf-8") as f:
        return yaml.safe_load(f)


@dataclass
class LoraConfig:
    """The subset of peft.LoraConfig the reference sets (models/lora_adapter.py:35-42)."""
    r: int = 8
    lora_alpha: int = 16
    lora_dropout: float = 0.1
    bias: str = "none"
    target_modules: List[str] = field(default_factory=lambda: ["q_proj", "v_proj"])
    task_type: str = "FEATURE_EXTRACTION"

    @property
    def scaling(self) -> float:
        return self.lora_alpha / self.r


def create_lora_config(config_path: Union[str, Path] = "config/lora_config.yaml") -> LoraConfig:
    """YAML -> LoraConfig with the reference's keys and defaults (r=8, alpha=16, dropout=0.1,
    bias="none", targets q_proj+v_proj, task FEATURE_EXTRACTION)."""
    cfg = _load_lora_config(config_path)
    lora_cfg = cfg.get("lora", {}) or {}
    model_cfg = cfg.get("model", {}) or {}
    target_modules = model_cfg.get("target_modules", ["q_proj", "v_proj"])
    return LoraConfig(
        r=lora_cfg.get("r", 8),
        lora_alpha=lora_cfg.get("alpha", 16),
        lora_dropout=lora_cfg.get("dropout", 0.1),
        bias=lora_cfg.get("bias", "none"),
        target_modules=list(target_modules),
        task_type=lora_cfg.get("task_type", "FEATURE_EXTRACTION"),
    )


# ------------------------------------------------------------------------------------------
# adapter container
# ------------------------------------------------------------------------------------------
@dataclass
class LoraAdapter:
    """Unmerged LoRA weights: module path -> (A [r, in], B [out, r]) fp32 CPU tensors.

    Module paths are relative to the CLIP model, e.g.
    ``vision_model.encoder.layers.0.self_attn.q_proj``.  y = base(x) + (x A^T) B^T * alpha/r.
    """
    config: LoraConfig
    weights: Dict[str, Tuple[torch.Tensor, torch.Tensor]]
    base_model_name_or_path: str = ""

    @property
    def scaling(self) -> float:
        return self.config.scaling

    def num_parameters(self) -> int:
        return sum(a.numel() + b.numel() for a, b in self.weights.values())


def linear_module_paths(vision_layers: int, text_layers: int) -> List[str]:
    """Dotted names of every nn.Linear PEFT would see inside a transformers CLIPModel."""
    names: List[str] = []
    for tower, n in (("text_model", text_layers), ("vision_model", vision_layers)):
        for i in range(n):
            base = f"{tower}.encoder.layers.{i}"
            for m in ("k_proj", "v_proj", "q_proj", "out_proj"):
                names.append(f"{base}.self_attn.{m}")
            names.append(f"{base}.mlp.fc1")
            names.append(f"{base}.mlp.fc2")
    names += ["visual_projection", "text_projection"]
    return names


def match_target_modules(names: Iterable[str], target_modules: Iterable[str]) -> List[str]:
    """PEFT's rule for a list of targets: wrap a module iff its name equals a target or ends with
    '.' + target.  Hence 'q_proj' hits BOTH towers."""
    targets = list(target_modules)
    return [n for n in names if any(n == t or n.endswith("." + t) for t in targets)]


def _check_supported(paths: Iterable[str]) -> None:
    for p in paths:
        leaf = p.rsplit(".", 1)[-1]
        if leaf not in SUPPORTED_TARGETS:
            raise NotImplementedError(
                f"LoRA on '{p}' is not supported by the B200 path; supported targets are "
                f"{SUPPORTED_TARGETS} (every Linear of the encoder layers; the two projection heads "
                f"visual_projection / text_projection are not LoRA targets here)")


def init_lora_adapter(model_dims: Dict[str, Tuple[int, int]], lora_config: LoraConfig,
                      seed: Optional[int] = None, init_b_std: float = 0.0,
                      base_model_name: str = "") -> LoraAdapter:
    """Fresh adapter the way get_peft_model creates it: A ~ kaiming_uniform(a=sqrt(5)), B = 0.
    `init_b_std` > 0 gives B ~ N(0, std^2) instead (tests need a LoRA that is not a no-op)."""
    paths = match_target_modules(model_dims.keys(), lora_config.target_modules)
    _check_supported(paths)
    gen = torch.Generator(device="cpu")
    if seed is not None:
        gen.manual_seed(seed)
    weights: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
    for p in paths:
        out_f, in_f = model_dims[p]
        a = torch.empty((lora_config.r, in_f), dtype=torch.float32)
        bound = 1.0 / math.sqrt(in_f)  # kaiming_uniform_(a=sqrt(5)) on a [r, in] matrix
        a.uniform_(-bound, bound, generator=gen)
        if init_b_std > 0:
            b = torch.randn((out_f, lora_config.r), generator=gen) * init_b_std
        else:
            b = torch.zeros((out_f, lora_config.r), dtype=torch.float32)
        weights[p] = (a, b)
    return LoraAdapter(config=lora_config, weights=weights, base_model_name_or_path=base_model_name)


def attach_lora_to_clip(model, lora_config: LoraConfig):
    """Attach a freshly initialised LoRA adapter to a B200ClipModel (reference signature).

    Like peft.get_peft_model + print_trainable_parameters (reference lines 53-55): every
    matching Linear in both towers gets an adapter (A random, B zero => output unchanged until
    trained/loaded) and the trainable-parameter count is printed."""
    adapter = init_lora_adapter(model.linear_dims(), lora_config,
                                base_model_name=getattr(model, "name", ""))
    model.set_lora(adapter)
    trainable = adapter.num_parameters()
    total = model.num_base_parameters() + trainable
    print(f"trainable params: {trainable:,d} || all params: {total:,d} || "
          f"trainable%: {100 * trainable / total:.4f}")
    return model


# ------------------------------------------------------------------------------------------
# PEFT on-disk layout
# ------------------------------------------------------------------------------------------
_PREFIX = "base_model.model."


def save_lora_adapter(adapter: LoraAdapter, out_dir: Union[str, Path]) -> Path:
    """Write adapter_config.json + adapter_model.safetensors exactly as peft's save_pretrained
    lays them out (keys `base_model.model.<path>.lora_{A,B}.weight`)."""
    from safetensors.torch import save_file

    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    c = adapter.config
    cfg = {
        "peft_type": "LORA",
        "task_type": c.task_type,
        "base_model_name_or_path": adapter.base_model_name_or_path,
        "r": c.r,
        "lora_alpha": c.lora_alpha,
        "lora_dropout": c.lora_dropout,
        "bias": c.bias,
        "target_modules": list(c.target_modules),
        "inference_mode": True,
        "fan_in_fan_out": False,
        "modules_to_save": None,
        "init_lora_weights": True,
    }
    (out / ADAPTER_CONFIG).write_text(json.dumps(cfg, indent=2, sort_keys=True))
    tensors = {}
    for path, (a, b) in adapter.weights.items():
        tensors[f"{_PREFIX}{path}.lora_A.weight"] = a.detach().to(torch.float32).contiguous().cpu()
        tensors[f"{_PREFIX}{path}.lora_B.weight"] = b.detach().to(torch.float32).contiguous().cpu()
    save_file(tensors, str(out / ADAPTER_WEIGHTS))
    return out


def load_lora_adapter(adapter_dir: Union[str, Path]) -> LoraAdapter:
    """Read a PEFT adapter directory (what PeftModel.from_pretrained consumes)."""
    from safetensors.torch import load_file

    d = Path(adapter_dir)
    cfg_path = d / ADAPTER_CONFIG
    w_path = d / ADAPTER_WEIGHTS
    if not cfg_path.exists():
        raise FileNotFoundError(f"LoRA adapter config not found: {cfg_path}")
    if not w_path.exists():
        bin_path = d / "adapter_model.bin"
        if not bin_path.exists():
            raise FileNotFoundError(f"LoRA adapter weights not found: {w_path}")
        tensors = torch.load(bin_path, map_location="cpu")
    else:
        tensors = load_file(str(w_path))
    raw = json.loads(cfg_path.read_text())
    if raw.get("peft_type", "LORA") != "LORA":
        raise ValueError(f"unsupported peft_type {raw.get('peft_type')!r}")
    targets = raw.get("target_modules") or ["q_proj", "v_proj"]
    if isinstance(targets, str):
        targets = [targets]
    cfg = LoraConfig(r=int(raw.get("r", 8)), lora_alpha=raw.get("lora_alpha", 16),
                     lora_dropout=raw.get("lora_dropout", 0.0), bias=raw.get("bias", "none"),
                     target_modules=list(targets),
                     task_type=raw.get("task_type") or "FEATURE_EXTRACTION")
    pairs: Dict[str, Dict[str, torch.Tensor]] = {}
    for key, t in tensors.items():
        k = key[len(_PREFIX):] if key.startswith(_PREFIX) else key
        k = k.replace(".lora_A.default.weight", ".lora_A.weight").replace(
            ".lora_B.default.weight", ".lora_B.weight")
        if k.endswith(".lora_A.weight"):
            pairs.setdefault(k[: -len(".lora_A.weight")], {})["A"] = t.float()
        elif k.endswith(".lora_B.weight"):
            pairs.setdefault(k[: -len(".lora_B.weight")], {})["B"] = t.float()
    weights = {}
    for path, ab in pairs.items():
        if "A" not in ab or "B" not in ab:
            raise ValueError(f"adapter tensor pair incomplete for {path}")
        if ab["A"].shape[0] != cfg.r or ab["B"].shape[1] != cfg.r:
            raise ValueError(f"adapter rank mismatch for {path}: A {tuple(ab['A'].shape)} "
                             f"B {tuple(ab['B'].shape)} r={cfg.r}")
        weights[path] = (ab["A"], ab["B"])
    _check_supported(weights.keys())
    return LoraAdapter(config=cfg, weights=weights,
                       base_model_name_or_path=raw.get("base_model_name_or_path") or "")
