"""The bf16 residual stream (clm_tower_set_residual_dtype, include/clm_b200.h): the stream h between the layers
stored in bf16 instead of the reference's fp32 (models/clip_model.py:62-66 loads fp32 weights and runs fp32).

What is checked: the in-place bf16 TMA reduce-add epilogue (gemm_kernel<*, *, 4>) against fp32 arithmetic with the
SAME roundings, LayerNorm over a bf16 stream against torch, and whole encoders against the fp32 CPU oracle at the
north_star bar (cosine >= 0.999) plus the distance to this library's own fp32-stream result.
"""
import os

import pytest
import torch

from clip_lora_match_b200 import _lib, kernels as K
from oracle import clip_oracle as O

pytestmark = pytest.mark.gpu


def _randn(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def _variant():
    return _lib.load().clm_last_gemm_variant()


def _fp32_matmul(a, w):
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        return a.float() @ w.float().T
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def _check_inplace_bf16(name, got, h0, upd):
    """got must be bf16(h0 + bf16(upd)) up to one bf16 ulp of the update / of the result (accumulation order of
    the fp32 dot product can move either rounding by one step)."""
    upd_b = upd.bfloat16().float()
    ref = (h0.float() + upd_b).bfloat16().float()
    err = (got.float() - ref).abs()
    tol = 2.0 ** -7 * (ref.abs() + upd_b.abs()) + 1e-3
    bad = err > tol
    assert not bad.any(), (f"{name}: {int(bad.sum())}/{bad.numel()} off, max err {float(err.max()):.4g} "
                           f"at {bad.nonzero()[:3].tolist()}")
    # and it is an update, not a store: the mean distance to the plain product must be the residual's size
    assert float((got.float() - upd).abs().mean()) > 0.3 * float(h0.float().abs().mean())


@pytest.mark.parametrize("M,N,Kd,variant", [
    (16384, 768, 3072, 25624),      # fc2 of ViT-B/16 -> CTA-pair kernel
    (16384 - 100, 1024, 1024, 25624),  # out-proj of ViT-L/14, ragged last pair tile
    (394, 768, 768, None),          # single-CTA tiles (small M)
    (77, 512, 2048, None),
])
def test_gemm_inplace_bf16_residual(cuda_device, M, N, Kd, variant):
    d = cuda_device
    a = _randn((M, Kd), 60).bfloat16().to(d)
    w = _randn((N, Kd), 61, Kd ** -0.5).bfloat16().to(d)
    bias = _randn((N,), 62).to(d)
    guard = 64
    buf = _randn((M + guard, N), 63, 2.0).bfloat16().to(d)
    h0 = buf[:M].clone()
    tail0 = buf[M:].clone()
    h = buf[:M]
    out = K.gemm_epi(a, w, bias=bias, residual=h, out=h)
    v = _variant()
    assert v % 10 == 4, v
    if variant is not None:
        assert v == variant, v
    assert out.data_ptr() == h.data_ptr()
    _check_inplace_bf16(f"inplace_bf16_{M}x{N}x{Kd}", h, h0, _fp32_matmul(a, w) + bias)
    assert torch.equal(buf[M:], tail0), "rows past M were touched"


def test_gemm_inplace_bf16_residual_with_lora_extension(cuda_device):
    d = cuda_device
    M, D, cols = 8192, 768, 64
    x = _randn((M, D), 70).bfloat16().to(d)
    w = _randn((D, D), 71, D ** -0.5).bfloat16().to(d)
    bias = _randn((D,), 72).to(d)
    a_cat = _randn((cols, D), 73, D ** -0.5).bfloat16().to(d)
    b_cat = _randn((D, cols), 74, 0.1).bfloat16().to(d)
    t = K.gemm_epi(x, a_cat)
    h = _randn((M, D), 75, 2.0).bfloat16().to(d)
    h0 = h.clone()
    K.gemm_epi(x, w, bias=bias, a2=t, w2=b_cat, residual=h, out=h)
    assert _variant() == 25624, _variant()
    _check_inplace_bf16("inplace_bf16_lora", h, h0, _fp32_matmul(x, w) + _fp32_matmul(t, b_cat) + bias)


def test_bf16_residual_must_be_in_place(cuda_device):
    d = cuda_device
    a = torch.zeros((128, 64), dtype=torch.bfloat16, device=d)
    w = torch.zeros((64, 64), dtype=torch.bfloat16, device=d)
    r = torch.zeros((128, 64), dtype=torch.bfloat16, device=d)
    with pytest.raises(ValueError):
        K.gemm_epi(a, w, residual=r)
    with pytest.raises(ValueError):
        K.gemm_epi(a, w, residual=r, out=torch.empty_like(r))


@pytest.mark.parametrize("dim", [128, 512, 768, 1024])
@pytest.mark.parametrize("rows", [1, 37, 1000])
def test_layernorm_over_a_bf16_stream(cuda_device, dim, rows):
    x = (_randn((rows, dim), 1, 3.0) + 0.5).bfloat16()
    g = _randn((dim,), 2) * 0.2 + 1.0
    b = _randn((dim,), 3) * 0.1
    y = K.layernorm(x.to(cuda_device), g.to(cuda_device), b.to(cuda_device), 1e-5)
    ref = torch.nn.functional.layer_norm(x.float(), (dim,), g, b, 1e-5)
    assert y.dtype == torch.bfloat16
    assert torch.allclose(y.float().cpu(), ref, atol=2e-2, rtol=8e-3)


def _ln_ref(h, g, b, eps=1e-5):
    return torch.nn.functional.layer_norm(h.float(), (h.shape[1],), g, b, eps)


@pytest.mark.parametrize("dim", [128, 512, 768, 1024])
def test_row_stats(cuda_device, dim):
    h = (_randn((1000, dim), 5, 3.0) + 0.7).bfloat16()
    st = K.row_stats(h.to(cuda_device), 1e-5).cpu()
    hf = h.float()
    assert torch.allclose(st[:, 0], hf.mean(dim=1), atol=1e-5, rtol=1e-5)
    assert torch.allclose(st[:, 1], torch.rsqrt(hf.var(dim=1, unbiased=False) + 1e-5), atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("M,N,Kd,act,variant", [
    (8192, 3072, 768, K.EPI_QUICKGELU, 25621),   # fc1 of ViT-B/16 on the CTA-pair kernel
    (8192 - 77, 3072, 1024, K.EPI_NONE, 25621),  # QKV of ViT-L/14, ragged last pair tile
    (394, 2304, 768, K.EPI_NONE, None),          # single-CTA tiles
    (77, 2048, 512, K.EPI_QUICKGELU, None),
])
def test_gemm_with_folded_layernorm(cuda_device, M, N, Kd, act, variant):
    """act(LN(h) W^T + b) with gamma folded into W and (mean, rstd) applied in the epilogue, against fp32 torch on the
    same bf16 stream.  The stream has a per-row offset several times its spread: the rank-1 correction has to cancel it."""
    d = cuda_device
    h = (_randn((M, Kd), 80, 1.5) + _randn((M, 1), 81, 4.0)).bfloat16()
    w = _randn((N, Kd), 82, Kd ** -0.5)
    g = _randn((Kd,), 83) * 0.2 + 1.0
    b = _randn((Kd,), 84) * 0.1
    bias = _randn((N,), 85)
    wg, cs, bf = K.fold_layernorm(w, g, b, bias)
    hd = h.to(d)
    st = K.row_stats(hd, 1e-5)
    out = K.gemm_ln_epi(hd, wg.to(d), st, cs.to(d), bf.to(d), ln_mode=1, act=act)
    if variant is not None:
        assert _variant() == variant, _variant()
    ref = _ln_ref(hd, g.to(d), b.to(d)) @ w.to(d).T + bias.to(d)
    if act == K.EPI_QUICKGELU:
        ref = ref * torch.sigmoid(1.702 * ref)
    err = (out.float() - ref).abs()
    tol = 3e-2 + 1e-2 * ref.abs()   # bf16 output + bf16 rounding of W diag(gamma) over K terms
    assert not (err > tol).any(), f"max err {float(err.max()):.4g}, {int((err > tol).sum())} off"
    # fp32 output: the only error left is the rounding of the folded weight
    out32 = K.gemm_ln_epi(hd, wg.to(d), st, cs.to(d), bf.to(d), ln_mode=1, act=act, out_dtype=torch.float32)
    ref_w = _ln_ref(hd, torch.ones_like(g).to(d), torch.zeros_like(b).to(d)) @ wg.to(d).float().T + bf.to(d)
    if act == K.EPI_QUICKGELU:
        ref_w = ref_w * torch.sigmoid(1.702 * ref_w)
    assert torch.allclose(out32, ref_w, atol=4e-3, rtol=2e-3), float((out32 - ref_w).abs().max())


def test_gemm_with_folded_layernorm_and_lora_extension(cuda_device):
    """The LoRA pair in folded form: u = (LN(h) A^T) / rstd from ln_mode 2, then u (sB)^T inside the accumulator of
    the ln_mode 1 GEMM, whose epilogue multiplies by rstd."""
    d = cuda_device
    M, D, cols, r = 8192, 768, 64, 16
    h = (_randn((M, D), 90, 1.5) + _randn((M, 1), 91, 3.0)).bfloat16().to(d)
    w = _randn((3 * D, D), 92, D ** -0.5)
    g = _randn((D,), 93) * 0.2 + 1.0
    b = _randn((D,), 94) * 0.1
    bias = _randn((3 * D,), 95)
    a_cat = torch.zeros((cols, D)); a_cat[:2 * r] = _randn((2 * r, D), 96, D ** -0.5)
    b_cat = torch.zeros((3 * D, cols)); b_cat[:D, :r] = _randn((D, r), 97, 0.1); b_cat[2 * D:, r:2 * r] = _randn((D, r), 98, 0.1)
    b_cat_b = b_cat.bfloat16().to(d)
    st = K.row_stats(h, 1e-5)
    ag, s_a, c_a = K.fold_layernorm(a_cat, g, b)          # c_a = A beta
    u = K.gemm_ln_epi(h, ag.to(d), st, s_a.to(d), None, ln_mode=2)
    x = _ln_ref(h, g.to(d), b.to(d))
    t_ref = x @ a_cat.to(d).T
    t_got = u.float() * st[:, 1:2] + c_a.to(d)
    assert torch.allclose(t_got, t_ref, atol=3e-2, rtol=2e-2), float((t_got - t_ref).abs().max())
    # the adapter's constant (A beta)(sB)^T goes into the folded bias of the wide GEMM
    wg, cs, bf = K.fold_layernorm(w, g, b, bias + b_cat_b.float().cpu() @ c_a)
    out = K.gemm_ln_epi(h, wg.to(d), st, cs.to(d), bf.to(d), ln_mode=1, a2=u, w2=b_cat_b)
    assert _variant() == 25621, _variant()
    ref = x @ w.to(d).T + bias.to(d) + t_ref @ b_cat_b.float().T
    err = (out.float() - ref).abs()
    tol = 3e-2 + 1e-2 * ref.abs()
    assert not (err > tol).any(), f"max err {float(err.max()):.4g}"
    base = x @ w.to(d).T + bias.to(d)
    assert float((ref[:, :D] - base[:, :D]).abs().max()) > 1e-2  # the adapter does something on q


def _model(arch_name, device, residual_dtype, r=16, alpha=32, targets=("q_proj", "v_proj")):
    from clip_lora_match_b200.models import clip_model as CM
    from clip_lora_match_b200.models.lora_adapter import LoraAdapter, LoraConfig

    model = O.build_model(arch_name, seed=0)
    weights = O.synthetic_lora(model, r, alpha, targets, seed=1)
    arch = (CM.arch_from_hf_config(O.hf_config(arch_name), arch_name) if arch_name == "tiny-test"
            else CM.arch_from_name(arch_name))
    lora = LoraAdapter(LoraConfig(r=r, lora_alpha=alpha, target_modules=list(targets)), weights)
    gpu = CM.B200ClipModel(arch, O.base_state_dict(model), lora=lora, device=device, residual_dtype=residual_dtype)
    return model, gpu


@pytest.mark.parametrize("arch,n,targets", [
    ("tiny-test", 6, ("q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2")),
    ("openai/clip-vit-base-patch32", 4, ("q_proj", "v_proj")),
    ("openai/clip-vit-large-patch14", 2, ("q_proj", "v_proj")),
])
def test_encoder_with_bf16_stream_vs_oracle_and_vs_fp32_stream(cuda_device, arch, n, targets):
    torch.set_num_threads(os.cpu_count() or 8)
    model, gpu = _model(arch, cuda_device, "bfloat16", targets=targets)
    assert gpu.residual_dtype == "bfloat16"
    for h in gpu._towers.values():
        assert _lib.load().clm_tower_residual_dtype(h) == _lib.OUT_BF16
    pv = O.synth_images(n, seed=2)
    ids, mask = O.synth_captions(n, seed=3)
    img16, txt16 = gpu.encode_images(pv).cpu(), gpu.encode_texts(ids).cpu()
    mi = O.parity_metrics(img16, O.encode_images(model, pv))
    mt = O.parity_metrics(txt16, O.encode_texts(model, ids, mask))
    print(f"[bf16 stream vs oracle] {arch}: image {mi} text {mt}")
    assert mi["cos_min"] >= 0.999 and mt["cos_min"] >= 0.999, (mi, mt)        # north_star
    assert mi["rel_l2_max"] <= 0.04 and mt["rel_l2_max"] <= 0.04, (mi, mt)
    # the same stream with standalone LayerNorm passes instead of the folded GEMMs
    gpu.set_ln_fold(False)
    img16n, txt16n = gpu.encode_images(pv).cpu(), gpu.encode_texts(ids).cpu()
    ni, nt = O.parity_metrics(img16n, O.encode_images(model, pv)), O.parity_metrics(txt16n, O.encode_texts(model, ids, mask))
    print(f"[bf16 stream, LayerNorm passes, vs oracle] {arch}: image {ni} text {nt}")
    assert ni["cos_min"] >= 0.999 and nt["cos_min"] >= 0.999, (ni, nt)
    assert not torch.equal(img16n, img16)
    gpu.set_ln_fold(True)
    assert torch.equal(gpu.encode_images(pv).cpu(), img16)
    gpu.set_residual_dtype("float32")
    for h in gpu._towers.values():
        assert _lib.load().clm_tower_residual_dtype(h) == _lib.OUT_F32
    img32, txt32 = gpu.encode_images(pv).cpu(), gpu.encode_texts(ids).cpu()
    di, dt = O.parity_metrics(img16, img32), O.parity_metrics(txt16, txt32)
    print(f"[bf16 stream vs fp32 stream] {arch}: image {di} text {dt}")
    assert di["cos_min"] >= 0.9995 and dt["cos_min"] >= 0.9995, (di, dt)
    assert not torch.equal(img16, img32)  # the switch does something
    # raw (un-normalised) features follow too
    gpu.set_residual_dtype("bf16")
    raw = gpu.encode_images(pv, normalize=False).cpu()
    assert O.parity_metrics(raw, O.encode_images(model, pv, normalize=False))["rel_l2_max"] <= 0.04


def test_bf16_stream_at_benchmark_batch(cuda_device):
    """ViT-L/14 at the batch bench.py measures: the CTA-pair reduce-add kernels on the real shapes, 8 sampled rows
    against the oracle."""
    torch.set_num_threads(os.cpu_count() or 8)
    batch = 512
    model, gpu = _model("openai/clip-vit-large-patch14", cuda_device, "bfloat16")
    g = torch.Generator().manual_seed(2)
    pv = torch.randn((batch, 3, 224, 224), generator=g)
    got = gpu.encode_images(pv.to(cuda_device)).cpu()
    assert torch.isfinite(got).all()
    sel = torch.linspace(0, batch - 1, 8).long()
    m = O.parity_metrics(got[sel], O.encode_images(model, pv[sel]))
    print(f"[bf16 stream, L/14 batch 512] {m}")
    assert m["cos_min"] >= 0.999, m
    small = gpu.encode_images(pv[sel].to(cuda_device)).cpu()
    mb = O.parity_metrics(small, got[sel])
    assert mb["cos_min"] >= 0.9995, mb


def test_residual_dtype_argument_is_validated():
    from clip_lora_match_b200.models import clip_model as CM
    assert CM._residual_dtype_name("bf16") == "bfloat16"
    assert CM._residual_dtype_name(torch.float32) == "float32"
    with pytest.raises(ValueError):
        CM._residual_dtype_name("float16")
