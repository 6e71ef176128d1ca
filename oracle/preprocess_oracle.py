"""Python front end of the C preprocessing oracle (oracle/pil_resample.c) — TEST INFRASTRUCTURE."""
from __future__ import annotations

import numpy as np

from . import build_oracle

MEAN = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)  # config/clip_config.yaml:8-12
STD = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)


def clip_preprocess(arr: np.ndarray, s: int = 224):
    """uint8 RGB [H, W, 3] -> (pixel_values fp32 [3, s, s], uint8 crop [s, s, 3]) — reference
    models/clip_model.py:105-107 through the transformers-4.x (Pillow) processor."""
    lib = build_oracle.load()
    arr = np.ascontiguousarray(arr)
    h, w, _ = arr.shape
    pv = np.empty((3, s, s), dtype=np.float32)
    crop = np.empty((s, s, 3), dtype=np.uint8)
    rc = lib.clm_oracle_clip_preprocess(arr.ctypes.data, h, w, s, MEAN.ctypes.data, STD.ctypes.data,
                                        pv.ctypes.data, crop.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"clm_oracle_clip_preprocess failed with {rc}")
    return pv, crop
