"""Batched image embedding — B200 mirror of the reference's src/embedding/embed_image.py.

Same signatures and defaults: embed_image(model, processor, image, device, normalize)
(reference :22-54) and embed_images_batch(model, processor, images, device, normalize,
batch_size=16) (reference :57-98; returns torch.empty(0) for an empty list, :95-96).
The forward of every batch is one clm_encode_image call; the default batch_size is kept for
signature compatibility but large values (1024+) are what the B200 path is built for.
"""
from __future__ import annotations

from pathlib import Path
from typing import Union

import torch
from PIL import Image


def _load_image(image: Union[str, Path, Image.Image]) -> Image.Image:
    if isinstance(image, Image.Image):
        return image.convert("RGB")
    image_path = Path(image)
    if not image_path.exists():
        raise FileNotFoundError(f"Image not found: {image_path}")
    return Image.open(image_path).convert("RGB")


def embed_image(model, processor, image: Union[str, Path, Image.Image],
                device: Union[str, torch.device] = "cuda", normalize: bool = True) -> torch.Tensor:
    """One image -> (d,) CPU fp32."""
    model.eval()
    inputs = processor(images=_load_image(image), return_tensors="pt")
    with torch.no_grad():
        feats = model.encode_images(inputs["pixel_values"], normalize=normalize)
    return feats.squeeze(0).detach().cpu()


def embed_images_batch(model, processor, images: list, device: Union[str, torch.device] = "cuda",
                       normalize: bool = True, batch_size: int = 16) -> torch.Tensor:
    """List of images -> (N, d) CPU fp32."""
    model.eval()
    all_embeddings = []
    for i in range(0, len(images), batch_size):
        batch_imgs = [_load_image(im) for im in images[i:i + batch_size]]
        inputs = processor(images=batch_imgs, return_tensors="pt", padding=True)
        with torch.no_grad():
            feats = model.encode_images(inputs["pixel_values"], normalize=normalize)
        all_embeddings.append(feats.cpu())
    if not all_embeddings:
        return torch.empty(0)
    return torch.cat(all_embeddings, dim=0)
