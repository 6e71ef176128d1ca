"""LoRA fine-tuning of CLIP — B200 mirror of the reference's scripts/train_lora.py.

Same YAML (`config/lora_config.yaml`: model / lora / data / training sections), same loop (symmetric InfoNCE at
temperature 0.07, AdamW, linear warm-up / linear decay, gradient accumulation, clip_grad_norm_, a validation pass
and a PEFT-layout checkpoint per epoch: reference :111-249), same printed lines.  What runs underneath is
models/lora_trainer.LoraTrainer: forward, backward and optimizer are the sm_100a kernels behind the C-ABI
(include/clm_b200.h, "LoRA training step"); there is no autograd and no CPU fallback.

Differences, all stated in DESIGN.md §4.5: LoRA dropout (lora.dropout) is not applied; image augmentation
(reference src/preprocessing/augment.py) is the caller's business (`augmenter=`); with no checkpoint of
`base_model_name` on disk the base weights are random-init of that architecture (there is no network here).
"""
from __future__ import annotations

import argparse
import math
import os
import random
from pathlib import Path
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
import yaml
from torch.utils.data import DataLoader, Dataset

from .. import _lib
from ..models import clip_model as CM
from ..models.lora_adapter import attach_lora_to_clip, create_lora_config, save_lora_adapter
from ..models.lora_trainer import LoraTrainer


def set_seed(seed: int) -> None:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def load_lora_training_config(config_path: str | Path = "config/lora_config.yaml") -> Dict:
    path = Path(config_path)
    if not path.exists():
        raise FileNotFoundError(f"LoRA config file not found: {path}")
    with open(path, "r", encoding="utf-8") as f:
        return yaml.safe_load(f)


class ClipPairDataset(Dataset):
    """CSV with `image_path` and `text` columns -> {pixel_values (3,H,W), input_ids (L,), attention_mask (L,)} with
    the text padded to the tokenizer's max length (reference datasets/dataset.py:16-89 and
    src/preprocessing/clip_preprocess.py:63-80)."""

    def __init__(self, csv_path: str | Path, image_root_dir: str | Path = ".", processor=None,
                 use_augmentation: bool = False, augmenter: Optional[Callable] = None,
                 model_name: str = "openai/clip-vit-base-patch32") -> None:
        import pandas as pd

        self.csv_path = Path(csv_path)
        if not self.csv_path.exists():
            raise FileNotFoundError(f"CSV not found: {self.csv_path}")
        self.df = pd.read_csv(self.csv_path)
        if "image_path" not in self.df.columns or "text" not in self.df.columns:
            raise ValueError("CSV must contain 'image_path' and 'text' columns.")
        self.image_root_dir = Path(image_root_dir)
        self.augmenter = augmenter if use_augmentation else None
        self.processor = processor or CM.ClmProcessor(model_name)
        self.max_length = getattr(self.processor.tokenizer, "model_max_length", 77)

    def __len__(self) -> int:
        return len(self.df)

    def __getitem__(self, idx: int) -> dict:
        from PIL import Image

        row = self.df.iloc[idx]
        path = Path(row["image_path"])
        if not path.is_absolute():
            path = self.image_root_dir / path
        if not path.exists():
            raise FileNotFoundError(f"Image file not found: {path}")
        img = Image.open(path).convert("RGB")
        if self.augmenter is not None:
            img = self.augmenter(img)
        text = str(row["text"])
        inputs = self.processor(text=[text], images=img, return_tensors="pt", padding="max_length",
                                truncation=True, max_length=self.max_length)
        return {"pixel_values": inputs["pixel_values"].squeeze(0), "input_ids": inputs["input_ids"].squeeze(0),
                "attention_mask": inputs["attention_mask"].squeeze(0), "image_path": str(path), "raw_text": text}


def build_dataloaders(cfg: Dict, processor=None) -> Tuple[DataLoader, DataLoader]:
    data_cfg = cfg.get("data", {}) or {}
    train_csv, val_csv = data_cfg.get("train_csv"), data_cfg.get("val_csv")
    if train_csv is None or val_csv is None:
        raise ValueError("Please set 'train_csv' and 'val_csv' in lora_config.yaml under 'data'.")
    root = data_cfg.get("image_root_dir", ".")
    name = (cfg.get("model", {}) or {}).get("base_model_name", "openai/clip-vit-base-patch32")
    train_cfg = cfg.get("training", {}) or {}
    bs, nw = int(train_cfg.get("batch_size", 32)), int(train_cfg.get("num_workers", 4))
    train_ds = ClipPairDataset(train_csv, root, processor, model_name=name)
    val_ds = ClipPairDataset(val_csv, root, processor, model_name=name)
    return (DataLoader(train_ds, batch_size=bs, shuffle=True, num_workers=nw, pin_memory=True),
            DataLoader(val_ds, batch_size=bs, shuffle=False, num_workers=nw, pin_memory=True))


def compute_clip_contrastive_loss(image_features: torch.Tensor, text_features: torch.Tensor,
                                  temperature: float = 0.07) -> torch.Tensor:
    """Symmetric image<->text InfoNCE on un-normalised features (N, d) -> scalar tensor on the GPU
    (reference :83-108), through clm_clip_loss (fp32)."""
    if not image_features.is_cuda or not text_features.is_cuda:
        raise ValueError("features must be CUDA tensors (the B200 path has no CPU fallback)")
    if image_features.shape != text_features.shape or image_features.dim() != 2:
        raise ValueError(f"features must both be (N, d); got {tuple(image_features.shape)} and {tuple(text_features.shape)}")
    lib = _lib.load()
    fi = image_features.detach().to(torch.float32).contiguous()
    ft = text_features.detach().to(torch.float32).contiguous()
    n, d = fi.shape
    nb = lib.clm_clip_loss_workspace_bytes(n, d)
    ws = torch.empty(nb, dtype=torch.uint8, device=fi.device)
    loss = torch.zeros((), dtype=torch.float32, device=fi.device)
    _lib.check(lib.clm_clip_loss(fi.data_ptr(), ft.data_ptr(), n, d, float(temperature), 1.0, loss.data_ptr(), None, None,
                                 None, None, ws.data_ptr(), nb, _lib.cur_stream()), "clm_clip_loss")
    return loss


def lr_lambda(step: int, total_steps: int, warmup_steps: int) -> float:
    """Linear warm-up then linear decay to zero (the closure of reference :148-151)."""
    if step < warmup_steps:
        return float(step) / max(1, warmup_steps)
    return max(0.0, float(total_steps - step) / max(1, total_steps - warmup_steps))


def load_base_model(model_name: str, device: Optional[str] = None, seed: int = 0) -> CM.B200ClipModel:
    """CLIPModel.from_pretrained(model_name) of reference :124 as a B200ClipModel: a local checkpoint if there is
    one (errors are raised), random init of the named architecture otherwise."""
    dev = CM._get_device(device)
    ckpt = None if os.environ.get("CLM_RANDOM_INIT") == "1" else CM._find_local_checkpoint(model_name)
    if ckpt is None:
        arch = CM.arch_from_name(model_name)
        print(f"[train_lora] no local checkpoint for '{model_name}': random-init weights of that architecture (seed={seed})")
        sd = CM.random_init_state_dict(arch, seed)
    else:
        from transformers import CLIPModel

        hf = CLIPModel.from_pretrained(ckpt, local_files_only=True)
        arch, sd = CM.arch_from_hf_config(hf.config, model_name), hf.state_dict()
    return CM.B200ClipModel(arch, sd, lora=None, device=dev)


def train(config_path: str | Path = "config/lora_config.yaml", loaders: Optional[Tuple] = None,
          model: Optional[CM.B200ClipModel] = None) -> LoraTrainer:
    """The reference's train() (:111-249).  `loaders` = (train_loader, val_loader) of dict batches replaces the CSV
    loaders (tests, synthetic data); `model` replaces from_pretrained."""
    cfg = load_lora_training_config(config_path)
    train_cfg = cfg.get("training", {}) or {}
    set_seed(int(train_cfg.get("seed", 42)))
    model_name = (cfg.get("model", {}) or {}).get("base_model_name", "openai/clip-vit-base-patch32")
    if model is None:
        print(f"[train_lora] Loading base CLIP model: {model_name}")
        model = load_base_model(model_name)
    device = model.device
    print(f"[train_lora] Using device: {device}")
    print("[train_lora] Creating LoRA config and attaching to model...")
    attach_lora_to_clip(model, create_lora_config(config_path))

    lr = float(train_cfg.get("learning_rate", 1e-4))
    weight_decay = float(train_cfg.get("weight_decay", 0.01))
    num_epochs = int(train_cfg.get("num_epochs", 5))
    grad_accum_steps = int(train_cfg.get("gradient_accumulation_steps", 1))
    max_grad_norm = float(train_cfg.get("max_grad_norm", 1.0))
    logging_steps = int(train_cfg.get("logging_steps", 50))
    temperature = float(train_cfg.get("temperature", 0.07))
    output_dir = Path(train_cfg.get("output_dir", "models/saved/clip-lora"))
    output_dir.mkdir(parents=True, exist_ok=True)

    trainer = LoraTrainer(model, lr=lr, weight_decay=weight_decay, max_grad_norm=max_grad_norm,
                          temperature=temperature, grad_accum_steps=grad_accum_steps)
    print(f"[train_lora] Number of trainable parameters: {trainer.num_trainable_parameters():,}")
    train_loader, val_loader = loaders if loaders is not None else build_dataloaders(cfg)

    total_steps = num_epochs * math.ceil(len(train_loader) / grad_accum_steps)
    warmup_steps = int(total_steps * train_cfg.get("warmup_ratio", 0.1))
    global_step = 0
    for epoch in range(num_epochs):
        running_loss = 0.0
        print(f"\n[train_lora] ===== Epoch {epoch + 1}/{num_epochs} =====")
        for step, batch in enumerate(train_loader, start=1):
            pv = batch["pixel_values"].to(device, non_blocking=True)
            ids = batch["input_ids"].to(device, non_blocking=True)
            mask = batch["attention_mask"].to(device, non_blocking=True)
            cur_lr = lr * lr_lambda(global_step, total_steps, warmup_steps)
            if grad_accum_steps == 1:
                loss = trainer.step(pv, ids, mask, lr=cur_lr)
                stepped = True
            else:
                loss = trainer.forward_backward(pv, ids, mask)
                stepped = step % grad_accum_steps == 0
                if stepped:
                    trainer.optimizer_step(lr=cur_lr)
            running_loss += loss.item()
            if stepped:
                global_step += 1
                if global_step % logging_steps == 0:
                    nxt = lr * lr_lambda(global_step, total_steps, warmup_steps)
                    print(f"[train_lora] Step {global_step}/{total_steps} LR={nxt:.2e} "
                          f"Loss={running_loss / logging_steps:.4f}")
                    running_loss = 0.0
        val_total, val_batches = 0.0, 0
        for batch in val_loader:
            val_total += trainer.eval_loss(batch["pixel_values"].to(device), batch["input_ids"].to(device),
                                           batch["attention_mask"].to(device)).item()
            val_batches += 1
        avg_val = val_total / val_batches if val_batches else float("nan")
        print(f"[train_lora] Epoch {epoch + 1} validation loss: {avg_val:.4f}")
        epoch_dir = output_dir / f"epoch_{epoch + 1}"
        print(f"[train_lora] Saving LoRA model to {epoch_dir} ...")
        save_lora_adapter(trainer.export_adapter(), epoch_dir)
    trainer.sync_model()
    print("[train_lora] Training finished.")
    return trainer


def main(argv=None):
    root = Path(__file__).resolve().parents[1]
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--config", type=Path, default=root / "config" / "lora_config.yaml")
    a = ap.parse_args(argv)
    train(a.config)


if __name__ == "__main__":
    main()
