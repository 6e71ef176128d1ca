"""GPU unit tests: every sm_100a kernel, called through the C-ABI, against a plain torch fp32
restatement of the same op on the same seeded inputs.  Tolerances are stated per test.

On failure the test dumps the offending tensors to gpurun_out/debug/ so one GPU round-trip
gives enough evidence to fix descriptor / layout mistakes.
"""
import os

import pytest
import torch

from clip_lora_match_b200 import kernels as K

pytestmark = pytest.mark.gpu

DUMP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "debug")


def _dump(name, **tensors):
    os.makedirs(DUMP, exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in tensors.items()}, os.path.join(DUMP, name + ".pt"))


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _randn(shape, seed, scale=1.0):
    return (torch.randn(shape, generator=_gen(seed)) * scale)


def _close(name, got, ref, atol, rtol, **extra):
    got = got.float().cpu()
    ref = ref.float().cpu()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    if bad.any() or not torch.isfinite(got).all():
        _dump(name, got=got, ref=ref, **extra)
        idx = bad.nonzero()[:5].tolist()
        raise AssertionError(
            f"{name}: {int(bad.sum())}/{bad.numel()} elements off; max err {float(err.max()):.4g}; "
            f"first bad {idx}; got {[float(got[tuple(i)]) for i in idx]} ref {[float(ref[tuple(i)]) for i in idx]}")


# ------------------------------------------------------------------------------------------
# elementwise
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [512, 768, 1024])
@pytest.mark.parametrize("rows", [1, 37, 1000])
def test_layernorm(cuda_device, dim, rows):
    x = _randn((rows, dim), 1, 3.0) + 0.5
    g = _randn((dim,), 2) * 0.2 + 1.0
    b = _randn((dim,), 3) * 0.1
    y = K.layernorm(x.to(cuda_device), g.to(cuda_device), b.to(cuda_device), 1e-5)
    ref = torch.nn.functional.layer_norm(x, (dim,), g, b, 1e-5)
    # output is bf16: half an ulp of bf16 (2^-9 relative) plus fp32 noise
    _close(f"layernorm_{rows}x{dim}", y, ref, atol=2e-3, rtol=4e-3)


@pytest.mark.parametrize("dim", [512, 768])
def test_l2norm(cuda_device, dim):
    x = _randn((33, dim), 4, 2.0)
    y, yb = K.l2norm(x.to(cuda_device), want_bf16=True)
    ref = x / x.norm(dim=-1, keepdim=True)
    _close(f"l2norm_{dim}", y, ref, atol=1e-6, rtol=1e-5)
    _close(f"l2norm_bf16_{dim}", yb, ref, atol=1e-4, rtol=4e-3)


# ------------------------------------------------------------------------------------------
# tcgen05 GEMM
# ------------------------------------------------------------------------------------------
def _gemm_ref(a, w, bias=None, residual=None, act=0, a2=None, w2=None):
    acc = a.float() @ w.float().T
    if a2 is not None:
        acc = acc + a2.float() @ w2.float().T
    if bias is not None:
        acc = acc + bias
    if act == K.EPI_QUICKGELU:
        acc = acc * torch.sigmoid(1.702 * acc)
    if residual is not None:
        acc = acc + residual
    return acc


GEMM_SHAPES = [
    # M, N, K           (what it covers)
    (128, 256, 64),      # one tile, one k-block
    (128, 256, 768),     # k loop across the 4-stage ring (phase flips)
    (256, 512, 128),     # 2x2 tiles
    (197 * 3, 2304, 768),  # ragged M (TMA zero fill + row guard), many tiles, QKV shape
    (1000, 768, 3072),   # fc2 shape, long K
    (300, 64, 768),      # BN=64 path (LoRA down-projection)
    (130, 128, 512),     # BN=128 path
    (77 * 5, 1536, 512),  # text tower QKV
    (50 * 4, 768, 588 + 52),  # patch-embed style K = 640 (588 padded)
]


@pytest.mark.parametrize("M,N,Kd", GEMM_SHAPES)
def test_gemm_plain(cuda_device, M, N, Kd):
    a = _randn((M, Kd), 10).bfloat16()
    w = _randn((N, Kd), 11, Kd ** -0.5).bfloat16()
    out = K.gemm_epi(a.to(cuda_device), w.to(cuda_device), out_dtype=torch.float32)
    ref = _gemm_ref(a, w)
    # bf16 products are exact in fp32; only the accumulation order differs
    _close(f"gemm_plain_{M}x{N}x{Kd}", out, ref, atol=2e-3, rtol=1e-3, a=a, w=w)


def test_gemm_identity_layout(cuda_device):
    """A = I (128x64 one-hot rows) makes the output a copy of W^T: pinpoints descriptor/swizzle errors."""
    M, N, Kd = 128, 256, 64
    a = torch.zeros((M, Kd))
    a[torch.arange(M), torch.arange(M) % Kd] = 1.0
    w = (torch.arange(N * Kd, dtype=torch.float32).reshape(N, Kd) % 251) / 16.0
    out = K.gemm_epi(a.bfloat16().to(cuda_device), w.bfloat16().to(cuda_device), out_dtype=torch.float32)
    ref = _gemm_ref(a.bfloat16(), w.bfloat16())
    _close("gemm_identity", out, ref, atol=1e-3, rtol=0, a=a, w=w)


@pytest.mark.parametrize("M,N,Kd", [(197 * 2, 768, 768), (300, 3072, 768),
                                    (300, 64, 768),    # BN=64: one bf16 slab drained by half the epilogue warps
                                    (130, 136, 512),   # ragged N inside a slab (TMA clips columns >= N)
                                    (77, 40, 128),     # N < one slab, M < one tile
                                    (1000, 1024, 256)])
def test_gemm_epilogues(cuda_device, M, N, Kd):
    a = _randn((M, Kd), 20).bfloat16()
    w = _randn((N, Kd), 21, Kd ** -0.5).bfloat16()
    bias = _randn((N,), 22)
    res = _randn((M, N), 23)
    d = cuda_device
    # bias, bf16 out
    out = K.gemm_epi(a.to(d), w.to(d), bias=bias.to(d))
    _close(f"gemm_bias_bf16_{N}", out, _gemm_ref(a, w, bias), atol=2e-2, rtol=8e-3)
    # bias + quickgelu, bf16 out
    out = K.gemm_epi(a.to(d), w.to(d), bias=bias.to(d), act=K.EPI_QUICKGELU)
    _close(f"gemm_gelu_{N}", out, _gemm_ref(a, w, bias, act=K.EPI_QUICKGELU), atol=2e-2, rtol=8e-3)
    # bias + residual, fp32 out, in place on the residual buffer
    res_d = res.to(d).clone()
    out = K.gemm_epi(a.to(d), w.to(d), bias=bias.to(d), residual=res_d, out=res_d)
    _close(f"gemm_residual_{N}", out, _gemm_ref(a, w, bias, res), atol=2e-3, rtol=1e-3)
    # residual in a different buffer than the output (per-thread epilogue path), fp32 and bf16 out
    out = K.gemm_epi(a.to(d), w.to(d), bias=bias.to(d), residual=res.to(d), out_dtype=torch.float32)
    _close(f"gemm_residual_sep_{N}", out, _gemm_ref(a, w, bias, res), atol=2e-3, rtol=1e-3)
    out = K.gemm_epi(a.to(d), w.to(d), bias=bias.to(d), residual=res.to(d), act=K.EPI_QUICKGELU)
    _close(f"gemm_residual_sep_bf16_{N}", out, _gemm_ref(a, w, bias, res, act=K.EPI_QUICKGELU), atol=3e-2, rtol=8e-3)
    # fp32 out without residual, with activation
    out = K.gemm_epi(a.to(d), w.to(d), bias=bias.to(d), act=K.EPI_QUICKGELU, out_dtype=torch.float32)
    _close(f"gemm_gelu_f32_{N}", out, _gemm_ref(a, w, bias, act=K.EPI_QUICKGELU), atol=3e-3, rtol=2e-3)
    # output written into a column slice of a wider buffer (ldo > N): neighbours must stay untouched
    wide = torch.full((M, N + 64), 7.0, dtype=torch.float32, device=d)
    K.gemm_epi(a.to(d), w.to(d), bias=bias.to(d), out=wide[:, 32:32 + N])
    _close(f"gemm_slice_{N}", wide[:, 32:32 + N], _gemm_ref(a, w, bias), atol=2e-3, rtol=1e-3)
    assert bool((wide[:, :32] == 7.0).all()) and bool((wide[:, 32 + N:] == 7.0).all())


def test_gemm_lora_extension(cuda_device):
    """y = x W^T + (x A^T)(s B)^T with the second product folded in as extra K blocks."""
    M, D, r = 197 * 2, 768, 16
    x = _randn((M, D), 30).bfloat16()
    w = _randn((3 * D, D), 31, D ** -0.5).bfloat16()
    bias = _randn((3 * D,), 32)
    a_cat = torch.zeros((64, D))
    a_cat[:r] = _randn((r, D), 33, D ** -0.5)       # q
    a_cat[r:2 * r] = _randn((r, D), 34, D ** -0.5)  # v
    b_cat = torch.zeros((3 * D, 64))
    b_cat[:D, :r] = _randn((D, r), 35, 0.05) * 2.0
    b_cat[2 * D:, r:2 * r] = _randn((D, r), 36, 0.05) * 2.0
    d = cuda_device
    a_cat_b, b_cat_b = a_cat.bfloat16(), b_cat.bfloat16()
    t = K.gemm_epi(x.to(d), a_cat_b.to(d))  # bf16 [M,64]
    _close("lora_down", t, _gemm_ref(x, a_cat_b), atol=2e-2, rtol=8e-3)
    out = K.gemm_epi(x.to(d), w.to(d), bias=bias.to(d), a2=t, w2=b_cat_b.to(d), out_dtype=torch.float32)
    ref = _gemm_ref(x, w, bias, a2=t.cpu(), w2=b_cat_b)
    _close("gemm_lora_ext", out, ref, atol=3e-3, rtol=1e-3)
    # and it must differ from the base GEMM where LoRA applies (q, v) but not on k
    base = _gemm_ref(x, w, bias)
    assert (ref[:, :D] - base[:, :D]).abs().max() > 1e-2
    assert (ref[:, D:2 * D] - base[:, D:2 * D]).abs().max() == 0


# ---- the kernels the benchmark actually runs: cta_group::2 pair tiles (256 x 256), every epilogue --------------
# A GEMM goes to the CTA-pair kernel when N % 256 == 0 and it has at least ~38 pair tiles (clm_gemm_launch);
# every shape above is smaller, so these are the fp32 comparisons of gemm_kernel<256, 2, *>.  The reference is
# torch fp32 on the GPU (TF32 off): the CPU would need minutes for these sizes.
def _variant():
    from clip_lora_match_b200 import _lib
    return _lib.load().clm_last_gemm_variant()


def _gemm_ref_gpu(a, w, bias=None, residual=None, act=0, a2=None, w2=None):
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        return _gemm_ref(a, w, bias, residual, act, a2, w2)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


PAIR = 25620  # BN 256, two CTAs; + epilogue id


@pytest.mark.parametrize("M", [8192, 197 * 50])  # full pair tiles / ragged last pair tile (9850 = 38 * 256 + 122)
def test_gemm_pair_bias_quickgelu_bf16(cuda_device, M):
    """(8192, 3072, 768) bias + QuickGELU, bf16 store: the vision fc1 launch -> gemm_kernel<256,2,1>."""
    d = cuda_device
    a = _randn((M, 768), 40).bfloat16().to(d)
    w = _randn((3072, 768), 41, 768 ** -0.5).bfloat16().to(d)
    bias = _randn((3072,), 42).to(d)
    out = K.gemm_epi(a, w, bias=bias, act=K.EPI_QUICKGELU)
    assert _variant() == PAIR + 1, _variant()
    _close(f"pair_gelu_{M}", out, _gemm_ref_gpu(a, w, bias, act=K.EPI_QUICKGELU), atol=2e-2, rtol=8e-3)


def test_gemm_pair_reduce_add_residual(cuda_device):
    """(16384, 768, 3072) h += g W2^T + b in place: the fc2 / out-proj launch -> gemm_kernel<256,2,3>
    (TMA reduce-add of fp32 tiles into the residual stream)."""
    d = cuda_device
    M, N, Kd = 16384, 768, 3072
    a = _randn((M, Kd), 43).bfloat16().to(d)
    w = _randn((N, Kd), 44, Kd ** -0.5).bfloat16().to(d)
    bias = _randn((N,), 45).to(d)
    res = _randn((M, N), 46).to(d)
    ref = _gemm_ref_gpu(a, w, bias, res)
    h = res.clone()
    out = K.gemm_epi(a, w, bias=bias, residual=h, out=h)
    assert _variant() == PAIR + 3, _variant()
    assert out.data_ptr() == h.data_ptr()
    _close("pair_reduce_add", h, ref, atol=3e-3, rtol=1e-3)
    # ragged M: rows past M of the last pair tile must not be touched (guard rows behind the buffer)
    M2 = 16384 - 100
    buf = res.clone()
    h2 = buf[:M2]
    K.gemm_epi(a[:M2], w, bias=bias, residual=h2, out=h2)
    assert _variant() == PAIR + 3
    _close("pair_reduce_add_ragged", h2, ref[:M2], atol=3e-3, rtol=1e-3)
    assert torch.equal(buf[M2:], res[M2:])


@pytest.mark.parametrize("cols", [64, 128])
def test_gemm_pair_lora_extension(cuda_device, cols):
    """(8192, 2304, 768) fused QKV with the LoRA K-extension (kb_ext = cols / 64 extra k-blocks read from a
    second pair of tensor maps) on the pair kernel -> gemm_kernel<256,2,1>."""
    d = cuda_device
    M, D = 8192, 768
    x = _randn((M, D), 50).bfloat16().to(d)
    w = _randn((3 * D, D), 51, D ** -0.5).bfloat16().to(d)
    bias = _randn((3 * D,), 52).to(d)
    a_cat = _randn((cols, D), 53, D ** -0.5).bfloat16().to(d)
    b_cat = (_randn((3 * D, cols), 54, 0.05) * 2.0)
    b_cat[D:2 * D] = 0  # no adapter on k
    b_cat = b_cat.bfloat16().to(d)
    t = K.gemm_epi(x, a_cat)
    _close(f"pair_lora_down_{cols}", t, _gemm_ref_gpu(x, a_cat), atol=2e-2, rtol=8e-3)
    out = K.gemm_epi(x, w, bias=bias, a2=t, w2=b_cat)
    assert _variant() == PAIR + 1, _variant()
    ref = _gemm_ref_gpu(x, w, bias, a2=t, w2=b_cat)
    _close(f"pair_lora_ext_{cols}", out, ref, atol=2e-2, rtol=8e-3)
    base = _gemm_ref_gpu(x, w, bias)
    assert (ref[:, :D] - base[:, :D]).abs().max() > 1e-2          # the extension is not a no-op on q ...
    assert (ref[:, D:2 * D] - base[:, D:2 * D]).abs().max() == 0  # ... and leaves k alone


def test_gemm_pair_store_f32(cuda_device):
    """fp32 TMA tile store on the pair kernel (the search's threshold-sampling GEMM) -> gemm_kernel<256,2,2>."""
    d = cuda_device
    a = _randn((8192, 1024), 55).bfloat16().to(d)
    w = _randn((2048, 1024), 56, 1024 ** -0.5).bfloat16().to(d)
    out = K.gemm_epi(a, w, out_dtype=torch.float32)
    assert _variant() == PAIR + 2, _variant()
    _close("pair_store_f32", out, _gemm_ref_gpu(a, w), atol=2e-3, rtol=1e-3)


# ------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------
def _attn_ref(qkv, batch, tokens, heads, causal):
    D = heads * 64
    q, k, v = qkv.float().reshape(batch, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    if causal:
        mask = torch.full((tokens, tokens), float("-inf")).triu(1)
        s = s + mask
    p = torch.softmax(s, dim=-1)
    o = p @ v
    return o.permute(0, 2, 1, 3).reshape(batch * tokens, D)


@pytest.mark.parametrize("batch,tokens,heads,causal", [
    (2, 64, 1, False),    # exactly one box, one head
    (3, 50, 12, False),   # ViT-B/32
    (2, 197, 12, False),  # ViT-B/16: two query tiles, ragged keys
    (2, 257, 16, False),  # ViT-L/14: S wider than one UMMA N (256 + 16)
    (5, 77, 8, True),     # text tower, causal
    (2, 77, 12, True),
    (3, 16, 2, False),    # one 16-key chunk: the second thread of every row has nothing to do
    (2, 130, 4, False),   # second query tile with 2 live rows; two S regions + two O regions
    (2, 225, 2, False),   # two key blocks of 128 + 112 keys (online softmax), per-thread output stores
    (1, 320, 2, True),    # causal and longer than 224: one S region, tiles strictly in sequence, one stage
    (1, 384, 1, False),   # the longest supported sequence (one pipeline stage -> one-region plan)
    (3, 129, 4, False),   # 128 + 1: extra-token path (CUDA-core key / query row) with two S + two O regions
    (40, 257, 8, False),  # ViT-L/14 extra-token path, 320 items on 148 SMs: stage reuse across items
    (37, 197, 12, False), # ViT-B/16, three items per CTA (stage / TMEM parity wrap-around)
    (7, 16, 8, True),     # length-bucketed text passes: causal at every bucket top
    (5, 32, 8, True),
    (4, 48, 12, True),
    (3, 64, 8, True),
    (300, 16, 8, True),   # many short captions: 2400 items
])
def test_attention(cuda_device, batch, tokens, heads, causal):
    D = heads * 64
    qkv = _randn((batch * tokens, 3 * D), 40, 1.5).bfloat16()
    out = K.attention(qkv.to(cuda_device), batch, tokens, heads, causal)
    ref = _attn_ref(qkv, batch, tokens, heads, causal)
    # P is rounded to bf16 before PV and the output is bf16
    _close(f"attention_{batch}x{tokens}x{heads}_{int(causal)}", out, ref, atol=1.5e-2, rtol=1e-2, qkv=qkv)


@pytest.mark.parametrize("mode", ["0", "2"])
@pytest.mark.parametrize("batch,tokens,heads", [
    (2, 257, 16),   # ViT-L/14: split kernel by default, alternate-chunk kernel with CLM_ATTN_SPLIT=0
    (40, 257, 8),   # several items per CTA: stage ring, TMEM region reuse, the four extra-token warps
    (37, 197, 12),  # ViT-B/16 (208 keys): split kernel by default as well, alternate-chunk kernel with CLM_ATTN_SPLIT=0
    (3, 200, 2),    # 208 keys with 8 masked ones
    (5, 193, 3),    # the smallest T of the 208-key plan
])
def test_attention_split_and_alternate_kernels(cuda_device, monkeypatch, mode, batch, tokens, heads):
    """Both softmax kernels for the ViT shapes (attention_kernel_split<Tk> and attention_kernel) against fp32:
    CLM_ATTN_SPLIT is read on every call (0 = never split, 1 / 2 / unset = split for 208 and 256 keys)."""
    monkeypatch.setenv("CLM_ATTN_SPLIT", mode)
    D = heads * 64
    qkv = _randn((batch * tokens, 3 * D), 41, 1.5).bfloat16()
    out = K.attention(qkv.to(cuda_device), batch, tokens, heads, False)
    ref = _attn_ref(qkv, batch, tokens, heads, False)
    _close(f"attention_split{mode}_{batch}x{tokens}x{heads}", out, ref, atol=1.5e-2, rtol=1e-2, qkv=qkv)


@pytest.mark.parametrize("mode,batch,tokens,heads", [("1", 300, 257, 16), ("2", 411, 197, 12)])
def test_attention_split_kernel_is_deterministic_with_many_items_per_cta(cuda_device, monkeypatch, mode, batch, tokens, heads):
    """30+ items per CTA (stage ring, TMEM regions, mbarrier phases wrap many times; the four cooperative
    extra-token warps and both MMA issuers run far ahead of / behind the softmax groups): the kernel has no atomics,
    so five runs must agree bit for bit -- a hand-off race shows up as a difference -- and sampled items must match
    the fp32 reference."""
    monkeypatch.setenv("CLM_ATTN_SPLIT", mode)
    D = heads * 64
    qkv = _randn((batch * tokens, 3 * D), 43, 1.5).bfloat16()
    qd = qkv.to(cuda_device)
    outs = [K.attention(qd, batch, tokens, heads, False) for _ in range(5)]
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    pick = [0, 1, batch // 2, batch - 2, batch - 1]
    sub = torch.cat([qkv[b * tokens:(b + 1) * tokens] for b in pick])
    ref = _attn_ref(sub, len(pick), tokens, heads, False)
    got = torch.cat([outs[0][b * tokens:(b + 1) * tokens] for b in pick]).float().cpu()
    assert torch.allclose(got, ref, atol=1.5e-2, rtol=1e-2), float((got - ref).abs().max())


@pytest.mark.parametrize("batch,tokens,heads", [(3, 257, 4), (3, 197, 4), (2, 200, 2), (2, 256, 2), (2, 145, 2)])
def test_attention_key_blocks_with_late_maximum(cuda_device, batch, tokens, heads):
    """The key-blocked kernel exponentiates block 1 (keys 128..) relative to the block-0 maximum and rescales
    O = P0 V0 only when a row's maximum moves up by more than 8 (log2 units).  Keys of the second block are
    scaled so that, per head, some rows need the rescale (logit gaps far above 8 / (0.125 log2 e) = 44), some
    do not, and in some the late keys are far BELOW the early ones (their P underflows to 0)."""
    D = heads * 64
    g = _gen(44)
    x = torch.randn((batch, tokens, 3, heads, 64), generator=g) * 1.2
    x[:, 128:, 1, 0] *= 6.0            # head 0: late keys dominate -> every row rescales
    x[:, 128:, 1, 1 % heads] *= 0.05   # head 1: late keys negligible
    x[:, 140:150, 1, heads - 1] *= 5.0  # last head: a few late outliers, only rows aligned with them rescale
    qkv = x.reshape(batch * tokens, 3 * D).bfloat16()
    out = K.attention(qkv.to(cuda_device), batch, tokens, heads, False)
    ref = _attn_ref(qkv, batch, tokens, heads, False)
    assert torch.isfinite(out).all()
    _close(f"attention_late_max_{batch}x{tokens}x{heads}", out, ref, atol=2e-2, rtol=1.5e-2, qkv=qkv)


# ------------------------------------------------------------------------------------------
# search
# ------------------------------------------------------------------------------------------
def _unit_rows(n, d, seed):
    x = _randn((n, d), seed)
    return x / x.norm(dim=-1, keepdim=True)


def _check_topk(name, got_s, got_i, q, e, k):
    """ids must match the fp32 oracle except for ties within 1e-4 in oracle score (north_star)."""
    sims = q @ e.T
    ref_s, ref_i = torch.topk(sims, k, dim=-1, largest=True, sorted=True)
    got_s, got_i = got_s.cpu(), got_i.cpu()
    assert got_i.shape == ref_i.shape
    assert (got_i >= 0).all() and (got_i < e.shape[0]).all(), f"{name}: ids out of range"
    if not torch.allclose(got_s, ref_s, atol=2e-6, rtol=1e-5):
        _dump(name, got_s=got_s, got_i=got_i, ref_s=ref_s, ref_i=ref_i)
        raise AssertionError(f"{name}: scores differ, max {float((got_s - ref_s).abs().max()):.3g}")
    mism = got_i != ref_i
    if mism.any():
        s_at_got = torch.gather(sims, 1, got_i)
        gap = (s_at_got - ref_s).abs()[mism]
        if not (gap <= 1e-4).all():
            _dump(name, got_s=got_s, got_i=got_i, ref_s=ref_s, ref_i=ref_i)
            raise AssertionError(f"{name}: {int(mism.sum())} id mismatches beyond the 1e-4 tie window")
    # no duplicates per query
    for row in got_i.tolist():
        assert len(set(row)) == len(row), f"{name}: duplicate ids"


@pytest.mark.parametrize("nq,n,d,k", [
    (1, 6, 512, 3),          # the shipped fixture's size: index smaller than a tile, k < n
    (1, 300, 512, 5),        # reference-style single query
    (16, 10000, 512, 10),    # config 1
    (1000, 10000, 512, 10),  # config 1, full query set
    (130, 70000, 768, 10),   # two query tiles, many splits, ragged last tile
    (64, 40000, 768, 50),    # k = 50 (config 5): 64-entry lists
    (300, 50000, 512, 100),  # top_k beyond one list: kc = 64 < k, no shared bound, CTA-pair scan
    (3, 9000, 512, 700),     # large k on the fused path (k <= 1024)
    (2, 3000, 256, 1500),    # beyond the fused path: exact per-query scan (clm_cosine_gemv + clm_topk_row)
])
def test_search_topk(cuda_device, nq, n, d, k):
    e = _unit_rows(n, d, 50)
    q = _unit_rows(nq, d, 51)
    dv = cuda_device
    kk = min(k, n)
    s, i = K.search_topk(q.to(dv), q.bfloat16().to(dv), e.bfloat16().to(dv), e.to(dv), kk)
    _check_topk(f"search_{nq}x{n}x{d}_k{k}", s, i, q, e, kk)


def test_search_id_offset_and_merge(cuda_device):
    """Row-sharded search: two shards searched separately then merged == unsharded search."""
    n, d, k, nq = 20000, 512, 10, 40
    e = _unit_rows(n, d, 60)
    q = _unit_rows(nq, d, 61)
    dv = cuda_device
    half = n // 2
    parts = []
    for lo, hi in ((0, half), (half, n)):
        es = e[lo:hi].contiguous()
        parts.append(K.search_topk(q.to(dv), q.bfloat16().to(dv), es.bfloat16().to(dv), es.to(dv), k,
                                   id_offset=lo))
    scores = torch.stack([p[0] for p in parts], dim=1).contiguous()
    ids = torch.stack([p[1] for p in parts], dim=1).contiguous()
    s, i = K.topk_merge_sorted(scores, ids, k)
    _check_topk("search_sharded_merge", s, i, q, e, k)


def test_kth_lower_bound_is_a_valid_and_tight_seed(cuda_device):
    """clm_kth_lower_bound (k-th largest of 1024 strided-group maxima) never exceeds the exact k-th largest of
    the row -- that is what makes it a valid seed of the scan's bound -- and for k << 1024 it is at least the
    (2k + 8)-th largest (tight enough to be useful); ties, -inf padding, ragged row lengths."""
    from clip_lora_match_b200 import _lib
    from clip_lora_match_b200._lib import check, cur_stream, ptr

    lib = _lib.load()
    g = _gen(71)
    for rows, n in ((33, 16384), (7, 5000), (5, 1024), (3, 40000)):
        x = torch.randn((rows, n), generator=g) * 0.04
        x[0, ::5] = x[0, 2]                      # ties
        x[1, 300:] = float("-inf")               # fewer than 1024 finite values
        xd = x.to(cuda_device)
        srt = torch.sort(x, dim=-1, descending=True).values
        for kth in (1, 10, 50, 256):
            out = torch.empty(rows, dtype=torch.float32, device=cuda_device)
            check(lib.clm_kth_lower_bound(ptr(xd), rows, n, kth, 0.0, ptr(out), cur_stream()), "clm_kth_lower_bound")
            o = out.cpu()
            assert (o <= srt[:, kth - 1]).all(), f"n={n} kth={kth}: not a lower bound"
            loose = srt[:, min(2 * kth + 8, n) - 1]
            assert (o[2:] >= loose[2:]).all(), f"n={n} kth={kth}: bound too loose"
    with pytest.raises(RuntimeError):
        check(lib.clm_kth_lower_bound(ptr(xd), 3, 40000, 300, 0.0, ptr(out), cur_stream()), "clm_kth_lower_bound")


def test_kth_largest_and_topk_row_match_torch(cuda_device):
    """The two selection primitives against torch on adversarial value sets: heavy ties, negatives, -inf padding."""
    from clip_lora_match_b200 import _lib
    from clip_lora_match_b200._lib import check, cur_stream, ptr

    lib = _lib.load()
    dv = cuda_device
    g = _gen(70)
    rows, n = 37, 16384
    x = torch.randn((rows, n), generator=g)
    x[:, ::7] = x[:, 3:4]                    # every 7th value of a row is the same number: massive ties
    x[5] = torch.round(x[5] * 2) / 2          # a row with ~10 distinct values
    x[6, 100:] = float("-inf")                # mostly empty
    xd = x.to(dv)
    for kth in (1, 2, 10, 50, 99, 1000, n):
        out = torch.empty(rows, dtype=torch.float32, device=dv)
        check(lib.clm_kth_largest(ptr(xd), rows, n, kth, 0.0, ptr(out), cur_stream()), "clm_kth_largest")
        ref = torch.topk(x, kth, dim=-1).values[:, -1]
        assert torch.equal(out.cpu(), ref), f"kth={kth}"
    n2 = 300_000
    y = torch.randn(n2, generator=g)
    y[1000:1400] = y[999]                    # 401 equal values straddling the cut for some k
    yd = y.to(dv)
    order = torch.argsort(y, descending=True, stable=True)
    for k in (1, 10, 64, 65, 1024, 2048):
        os_ = torch.empty(k, dtype=torch.float32, device=dv)
        oi = torch.empty(k, dtype=torch.int64, device=dv)
        check(lib.clm_topk_row(ptr(yd), n2, k, 7, ptr(os_), ptr(oi), cur_stream()), "clm_topk_row")
        ref_s = y[order[:k]]
        assert torch.equal(os_.cpu(), ref_s), f"k={k}: scores"
        got = oi.cpu() - 7
        assert torch.equal(y[got], ref_s) and len(set(got.tolist())) == k, f"k={k}: ids"
        # sorted by (score desc, id asc)
        same = os_.cpu()[1:] == os_.cpu()[:-1]
        assert bool((got[1:][same] > got[:-1][same]).all())


def test_topk_merge_gathered_reads_the_packed_chunks_in_place(cuda_device):
    """The exchange layout (pack_topk_chunk -> rank-major buffer -> clm_topk_merge_gathered) against the plain
    [Q, lists, k] merge, including a rank that holds fewer than k rows and an odd Q*k (padding to an even count)."""
    dv = cuda_device
    g = _gen(71)
    world, nq, k = 3, 7, 5   # nq * k = 35: odd
    scores = torch.sort(torch.randn((world, nq, k), generator=g), dim=-1, descending=True).values
    ids = torch.randint(0, 1_000_000, (world, nq, k), generator=g)
    chunks = []
    for r in range(world):
        kl = 2 if r == 1 else k  # rank 1's shard has only 2 rows
        chunks.append(K.pack_topk_chunk(scores[r, :, :kl].to(dv), ids[r, :, :kl].to(dv), nq, k, dv))
    scores[1, :, 2:] = float("-inf"); ids[1, :, 2:] = -1
    buf = torch.cat(chunks)
    s, i = K.topk_merge_gathered(buf, world, nq, k)
    flat_s = scores.permute(1, 0, 2).reshape(nq, -1)
    flat_i = ids.permute(1, 0, 2).reshape(nq, -1)
    ref_s, pos = torch.topk(flat_s, k, dim=-1)
    assert torch.equal(s.cpu(), ref_s)
    assert torch.equal(i.cpu(), torch.gather(flat_i, 1, pos))
    s2, i2 = K.topk_merge_sorted(scores.permute(1, 0, 2).contiguous().to(dv), ids.permute(1, 0, 2).contiguous().to(dv), k)
    assert torch.equal(s2, s) and torch.equal(i2, i)
