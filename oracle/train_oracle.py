"""CPU oracle of the LoRA training step — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see clip_oracle.py's header).

Restates reference scripts/train_lora.py in plain fp32 PyTorch with autograd:
  * `contrastive_loss`  = compute_clip_contrastive_loss, train_lora.py:83-108 (symmetric InfoNCE);
  * `lr_lambda`         = the linear warm-up / linear decay closure, train_lora.py:148-151;
  * `TrainOracle.step`  = the body of the step loop, train_lora.py:172-193: get_image_features /
    get_text_features of the LoRA-wrapped transformers CLIPModel, loss / grad_accum_steps, backward,
    clip_grad_norm_(max_grad_norm), torch.optim.AdamW(lr, weight_decay).step(), zero_grad.
The model is the real transformers CLIPModel (eager attention) with clip_oracle.LoraLinear wrappers (PEFT
semantics).  LoRA dropout is 0 here: the B200 step does not apply dropout either
(DESIGN.md §4.5), and a stochastic mask could not be compared anyway.

Pinning: `contrastive_loss` and `lr_lambda` are checked against the reference's OWN functions, imported from
/root/reference/scripts/train_lora.py by oracle/make_golden.py (tests/golden/train_golden.npz,
tests/test_oracle_golden.py).  The optimizer is torch's own AdamW, as in the reference.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch
import torch.nn.functional as F

from . import clip_oracle as O


def contrastive_loss(image_features: torch.Tensor, text_features: torch.Tensor, temperature: float = 0.07) -> torch.Tensor:
    """train_lora.py:95-106."""
    image_features = image_features / image_features.norm(dim=-1, keepdim=True)
    text_features = text_features / text_features.norm(dim=-1, keepdim=True)
    logits_per_image = (image_features @ text_features.T) / temperature
    logits_per_text = logits_per_image.T
    targets = torch.arange(image_features.shape[0], device=image_features.device)
    return (F.cross_entropy(logits_per_image, targets) + F.cross_entropy(logits_per_text, targets)) / 2.0


def lr_lambda(step: int, total_steps: int, warmup_steps: int) -> float:
    """train_lora.py:148-151."""
    if step < warmup_steps:
        return float(step) / max(1, warmup_steps)
    return max(0.0, float(total_steps - step) / max(1, total_steps - warmup_steps))


class TrainOracle:
    def __init__(self, model, lr: float = 1e-4, weight_decay: float = 0.01, max_grad_norm: float = 1.0,
                 temperature: float = 0.07, grad_accum_steps: int = 1):
        """`model`: a transformers CLIPModel already wrapped by clip_oracle.inject_lora / synthetic_lora."""
        self.model = model
        self.model.train()  # dropout modules are identities (p = 0); CLIP itself has attention_dropout = 0
        self.params = [p for n, p in model.named_parameters() if ".lora_A." in n or ".lora_B." in n]
        for n, p in model.named_parameters():
            p.requires_grad_(".lora_A." in n or ".lora_B." in n)
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay)
        self.max_grad_norm, self.temperature, self.grad_accum_steps = max_grad_norm, temperature, grad_accum_steps

    def features(self, pixel_values, input_ids, attention_mask=None) -> Tuple[torch.Tensor, torch.Tensor]:
        fi = O._pooled(self.model.get_image_features(pixel_values=pixel_values.float()))
        ft = O._pooled(self.model.get_text_features(input_ids=input_ids.long(), attention_mask=attention_mask))
        return fi, ft

    def forward_backward(self, pixel_values, input_ids, attention_mask=None) -> float:
        fi, ft = self.features(pixel_values, input_ids, attention_mask)
        loss = contrastive_loss(fi, ft, self.temperature) / self.grad_accum_steps
        loss.backward()
        return float(loss.item())

    def gradients(self) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
        out = {}
        for n, m in self.model.named_modules():
            if isinstance(m, O.LoraLinear):
                out[n] = (m.lora_A.weight.grad.detach().clone(), m.lora_B.weight.grad.detach().clone())
        return out

    def optimizer_step(self, lr: Optional[float] = None) -> float:
        if lr is not None:
            for g in self.opt.param_groups:
                g["lr"] = lr
        norm = torch.nn.utils.clip_grad_norm_(self.params, self.max_grad_norm)
        self.opt.step()
        self.opt.zero_grad()
        return float(norm)

    def step(self, pixel_values, input_ids, attention_mask=None, lr: Optional[float] = None) -> float:
        loss = self.forward_backward(pixel_values, input_ids, attention_mask)
        self.optimizer_step(lr)
        return loss

    @torch.no_grad()
    def eval_loss(self, pixel_values, input_ids, attention_mask=None) -> float:
        fi, ft = self.features(pixel_values, input_ids, attention_mask)
        return float(contrastive_loss(fi, ft, self.temperature).item())
