"""The C oracle of the image preprocessing (oracle/pil_resample.c) pinned against Pillow itself
(bit exact: it restates Pillow's Resample.c) and against the installed CLIPImageProcessor."""
import ctypes as C

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import build_oracle
from oracle.preprocess_oracle import MEAN, STD
from oracle.preprocess_oracle import clip_preprocess as oracle_preprocess

SIZES = [(300, 260), (260, 300), (224, 224), (225, 224), (768, 1024), (100, 80), (50, 333), (223, 500), (641, 479)]


def pil_crop(arr, s=224):
    """transformers-4.x slow processor on the uint8 stage: Pillow bicubic shortest-edge resize + centre crop."""
    h, w, _ = arr.shape
    nw, nh = (s, int(s * h / w)) if w <= h else (int(s * w / h), s)
    r = np.asarray(Image.fromarray(arr, "RGB").resize((nw, nh), resample=Image.BICUBIC))
    top, left = (nh - s) // 2, (nw - s) // 2
    return r[top:top + s, left:left + s]


@pytest.mark.parametrize("h,w", SIZES)
def test_uint8_stage_is_bit_exact_with_pillow(h, w):
    rs = np.random.RandomState(h * 1000 + w)
    for arr in (rs.randint(0, 256, size=(h, w, 3), dtype=np.uint8),
                (np.add.outer(np.arange(h), np.arange(w))[..., None] * np.array([1, 2, 3]) % 256).astype(np.uint8)):
        arr = np.ascontiguousarray(arr)
        _, crop = oracle_preprocess(arr)
        assert np.array_equal(crop, pil_crop(arr))


def test_float_stage_matches_the_4x_processor_arithmetic():
    arr = np.random.RandomState(3).randint(0, 256, size=(333, 500, 3), dtype=np.uint8)
    pv, crop = oracle_preprocess(arr)
    ref = ((crop.astype(np.float64) * (1 / 255)).astype(np.float32) - MEAN) / STD      # rescale (f64) then normalize (f32)
    assert np.array_equal(pv, ref.transpose(2, 0, 1))


def test_within_one_level_of_installed_clip_image_processor():
    """transformers 5.5 resamples with torchvision (not Pillow): its uint8 result may differ from
    Pillow's by one level on some pixels; the float arithmetic differs by an ulp."""
    from transformers import CLIPImageProcessor

    arr = np.random.RandomState(5).randint(0, 256, size=(300, 260, 3), dtype=np.uint8)
    pv, _ = oracle_preprocess(arr)
    hf = CLIPImageProcessor()(images=Image.fromarray(arr, "RGB"), return_tensors="pt")["pixel_values"][0].numpy()
    one_level = (1 / 255) / STD.min()
    assert np.abs(pv - hf).max() <= one_level * 1.001
    assert np.abs(pv - hf).mean() <= 0.1 * one_level


def test_resized_shape_rule():
    lib = build_oracle.load()
    oh, ow = C.c_int(), C.c_int()
    for (h, w), exp in [((300, 260), (258, 224)), ((260, 300), (224, 258)), ((224, 224), (224, 224)), ((50, 333), (224, 1491))]:
        lib.clm_oracle_resized_shape(h, w, 224, C.byref(oh), C.byref(ow))
        assert (oh.value, ow.value) == exp
