"""Minimal stand-in for the `peft` package — TEST INFRASTRUCTURE (see clip_oracle.py header).

`peft` is a dependency of the reference (requirements.txt:6, unpinned) that is not installed
in this image and cannot be fetched.  The reference's models/clip_model.py:12 and
models/lora_adapter.py:9 import it at module top, so to execute the reference's OWN code in
oracle/make_golden.py we register this module as `peft`.  It implements exactly the three
names the reference uses, with PEFT's documented LoRA semantics:

  LoraConfig(r, lora_alpha, lora_dropout, bias, target_modules, task_type)
  get_peft_model(model, config) -> wrapper with print_trainable_parameters()/save_pretrained()
  PeftModel.from_pretrained(model, path) -> wrapper with the adapter loaded

LoRA arithmetic itself is oracle.clip_oracle.LoraLinear.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional

import torch
import torch.nn as nn

from . import clip_oracle as O


@dataclass
class LoraConfig:
    r: int = 8
    lora_alpha: int = 8
    lora_dropout: float = 0.0
    bias: str = "none"
    target_modules: Optional[List[str]] = None
    task_type: Optional[str] = None
    base_model_name_or_path: str = ""


class PeftModel(nn.Module):
    """FEATURE_EXTRACTION-style wrapper: unknown attributes forward to the wrapped model."""

    def __init__(self, model: nn.Module, config: LoraConfig):
        super().__init__()
        self.base_model = model
        self.peft_config = config

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("base_model"), name)

    def forward(self, *a, **k):
        return self.base_model(*a, **k)

    def trainable_parameter_count(self):
        trainable = sum(p.numel() for p in self.parameters() if p.requires_grad)
        total = sum(p.numel() for p in self.parameters())
        return trainable, total

    def print_trainable_parameters(self):
        t, a = self.trainable_parameter_count()
        print(f"trainable params: {t:,d} || all params: {a:,d} || trainable%: {100 * t / a:.4f}")

    def save_pretrained(self, out_dir):
        from safetensors.torch import save_file

        out = Path(out_dir)
        out.mkdir(parents=True, exist_ok=True)
        c = self.peft_config
        (out / "adapter_config.json").write_text(json.dumps({
            "peft_type": "LORA", "task_type": c.task_type, "r": c.r, "lora_alpha": c.lora_alpha,
            "lora_dropout": c.lora_dropout, "bias": c.bias, "target_modules": list(c.target_modules or []),
            "base_model_name_or_path": c.base_model_name_or_path, "inference_mode": True}, indent=2))
        tensors = {}
        for path, (a, b) in O.get_lora_weights(self.base_model).items():
            tensors[f"base_model.model.{path}.lora_A.weight"] = a.contiguous()
            tensors[f"base_model.model.{path}.lora_B.weight"] = b.contiguous()
        save_file(tensors, str(out / "adapter_model.safetensors"))

    @classmethod
    def from_pretrained(cls, model: nn.Module, path: str):
        from safetensors.torch import load_file

        d = Path(path)
        raw = json.loads((d / "adapter_config.json").read_text())
        cfg = LoraConfig(r=raw["r"], lora_alpha=raw["lora_alpha"], lora_dropout=raw.get("lora_dropout", 0.0),
                         bias=raw.get("bias", "none"), target_modules=raw["target_modules"],
                         task_type=raw.get("task_type"))
        O.inject_lora(model, cfg.r, cfg.lora_alpha, cfg.target_modules, cfg.lora_dropout)
        tensors = load_file(str(d / "adapter_model.safetensors"))
        weights = {}
        for k, t in tensors.items():
            k = k[len("base_model.model."):]
            if k.endswith(".lora_A.weight"):
                weights.setdefault(k[:-len(".lora_A.weight")], [None, None])[0] = t
            elif k.endswith(".lora_B.weight"):
                weights.setdefault(k[:-len(".lora_B.weight")], [None, None])[1] = t
        O.set_lora_weights(model, {k: (a, b) for k, (a, b) in weights.items()})
        return cls(model, cfg)


def get_peft_model(model: nn.Module, config: LoraConfig) -> PeftModel:
    for p in model.parameters():
        p.requires_grad_(False)
    O.inject_lora(model, config.r, config.lora_alpha, config.target_modules or [], config.lora_dropout)
    for m in model.modules():
        if isinstance(m, O.LoraLinear):
            m.lora_A.weight.requires_grad_(True)
            m.lora_B.weight.requires_grad_(True)
    return PeftModel(model, config)
