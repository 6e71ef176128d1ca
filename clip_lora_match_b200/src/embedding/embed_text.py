"""Text embedding for a string or a list — B200 mirror of the reference's
src/embedding/embed_text.py (embed_text :11-60): str -> (d,), list -> (N, d), CPU fp32,
tokenizer called with padding=True / truncation=True / max_length=model_max_length.
"""
from __future__ import annotations

from typing import List, Union

import torch


def embed_text(model, processor, text: Union[str, List[str]], device: Union[str, torch.device] = "cuda",
               normalize: bool = True) -> torch.Tensor:
    model.eval()
    single_input = isinstance(text, str)
    texts = [text] if single_input else list(text)
    if not texts:
        return torch.empty(0)
    enc = processor.tokenizer(texts, padding=True, truncation=True,
                              max_length=processor.tokenizer.model_max_length, return_tensors="pt")
    input_ids, attention_mask = enc["input_ids"], enc["attention_mask"]
    # padded positions -> EOS so the first-EOS pooling row is the caption's own EOS
    input_ids = torch.where(attention_mask.bool(), input_ids, torch.full_like(input_ids, model.arch.eos_id))
    with torch.no_grad():
        feats = model.encode_texts(input_ids, normalize=normalize)
    feats = feats.cpu()
    return feats.squeeze(0) if single_input else feats
