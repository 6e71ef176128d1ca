"""GPU tests of the LoRA training step (reference scripts/train_lora.py:83-108,170-211).

Kernel level: every kernel of csrc/clm_train.cu against torch (autograd) on the same inputs.
Step level: LoraTrainer (forward + backward + clip + AdamW through the C-ABI) against oracle/train_oracle.py -- the
real transformers CLIPModel with PEFT-semantics LoRA wrappers under torch autograd and torch.optim.AdamW in fp32.
Tolerances are for bf16 tensor-core operands against an fp32 reference: loss within 2e-2 absolute, gradient
cosine >= 0.98 per factor with a non-negligible norm and >= 0.99 over all factors together.
"""
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle as O
from oracle import train_oracle as T

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _lib():
    from clip_lora_match_b200 import _lib as L
    return L, L.load()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


# ------------------------------------------------------------------------------------------------------------
# kernels
# ------------------------------------------------------------------------------------------------------------
def test_quickgelu_fwd_bwd(cuda_device):
    L, lib = _lib()
    g = torch.Generator().manual_seed(0)
    z = (torch.randn((777, 256), generator=g) * 2.5).to(cuda_device, torch.bfloat16)
    dg = torch.randn((777, 256), generator=g).to(cuda_device, torch.bfloat16)
    out = torch.empty_like(z)
    dz = torch.empty_like(z)
    L.check(lib.clm_quickgelu_fwd(z.data_ptr(), out.data_ptr(), z.numel(), _st()))
    L.check(lib.clm_quickgelu_bwd(dg.data_ptr(), z.data_ptr(), dz.data_ptr(), z.numel(), _st()))
    zf = z.float().requires_grad_(True)
    ref = zf * torch.sigmoid(1.702 * zf)
    ref.backward(dg.float())
    assert torch.allclose(out.float(), ref.detach(), rtol=2e-2, atol=2e-3)
    assert torch.allclose(dz.float(), zf.grad, rtol=2e-2, atol=2e-3)
    # in place (the trainer overwrites dg)
    L.check(lib.clm_quickgelu_bwd(dg.data_ptr(), z.data_ptr(), dg.data_ptr(), z.numel(), _st()))
    assert torch.equal(dg, dz)


@pytest.mark.parametrize("dim", [128, 512, 768, 1024])
@pytest.mark.parametrize("dy_f32", [False, True])
def test_layernorm_bwd(cuda_device, dim, dy_f32):
    L, lib = _lib()
    g = torch.Generator().manual_seed(dim)
    rows = 203
    x = (torch.randn((rows, dim), generator=g) * 1.7 + 0.3).to(cuda_device)
    gamma = (1.0 + 0.2 * torch.randn(dim, generator=g)).to(cuda_device)
    beta = (0.1 * torch.randn(dim, generator=g)).to(cuda_device)
    dy = torch.randn((rows, dim), generator=g).to(cuda_device)
    dy_in = dy if dy_f32 else dy.to(torch.bfloat16)
    xr = x.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (dim,), gamma, beta, 1e-5).backward(dy_in.float())
    dres0 = torch.randn((rows, dim), generator=g).to(cuda_device)
    for acc in (0, 1):
        dres = dres0.clone()
        shadow = torch.zeros((rows, dim), dtype=torch.bfloat16, device=cuda_device)
        L.check(lib.clm_layernorm_bwd(dy_in.data_ptr(), int(dy_f32), x.data_ptr(), gamma.data_ptr(), dres.data_ptr(),
                                      shadow.data_ptr(), rows, dim, 1e-5, acc, None, 0, 0, _st()))
        ref = xr.grad + (dres0 if acc else 0)
        assert torch.allclose(dres, ref, rtol=1e-4, atol=1e-4), (dres - ref).abs().max()
        assert torch.equal(shadow, dres.to(torch.bfloat16))
    # gather mode: item b works on row b * tokens + idx[b]
    tokens, batch = 7, rows // 7
    idx = torch.randint(0, tokens, (batch,), generator=g).to(cuda_device, torch.int32)
    dres = torch.zeros((rows, dim), device=cuda_device)
    shadow = torch.zeros((rows, dim), dtype=torch.bfloat16, device=cuda_device)
    L.check(lib.clm_layernorm_bwd(dy_in.data_ptr(), int(dy_f32), x.data_ptr(), gamma.data_ptr(), dres.data_ptr(),
                                  shadow.data_ptr(), batch, dim, 1e-5, 0, idx.data_ptr(), tokens, 1, _st()))
    sel = torch.arange(batch, device=cuda_device) * tokens + idx.long()
    xg = x[sel].clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xg, (dim,), gamma, beta, 1e-5).backward(dy_in[:batch].float())
    ref = torch.zeros_like(dres)
    ref[sel] = xg.grad
    assert torch.allclose(dres, ref, rtol=1e-4, atol=1e-4)
    assert torch.equal(shadow, dres.to(torch.bfloat16))


def test_transpose_and_cast(cuda_device):
    L, lib = _lib()
    g = torch.Generator().manual_seed(3)
    for rows, cols, f32 in [(50, 64, False), (1576, 2304, False), (333, 72, True)]:
        src = torch.randn((rows, cols), generator=g).to(cuda_device)
        src = src if f32 else src.to(torch.bfloat16)
        ld = (rows + 7) // 8 * 8
        dst = torch.zeros((cols, ld), dtype=torch.bfloat16, device=cuda_device)
        L.check(lib.clm_transpose_to_bf16(src.data_ptr(), int(f32), cols, 0, rows, cols, dst.data_ptr(), ld, 0, 1, 1.0, _st()))
        assert torch.equal(dst[:, :rows], src.t().to(torch.bfloat16))
        assert not dst[:, rows:].any()
    # batched with a scale (the bf16 copies of the LoRA masters)
    src = torch.randn((3, 40, 64), generator=g).to(cuda_device)
    dst = torch.zeros((3, 64, 40), dtype=torch.bfloat16, device=cuda_device)
    L.check(lib.clm_transpose_to_bf16(src.data_ptr(), 1, 64, 40 * 64, 40, 64, dst.data_ptr(), 40, 64 * 40, 3, 2.0, _st()))
    assert torch.equal(dst, (2.0 * src).transpose(1, 2).to(torch.bfloat16))
    flat = torch.randn(4096, generator=g).to(cuda_device)
    out = torch.zeros(4096, dtype=torch.bfloat16, device=cuda_device)
    L.check(lib.clm_cast_to_bf16(flat.data_ptr(), out.data_ptr(), 4096, 0.5, _st()))
    assert torch.equal(out, (0.5 * flat).to(torch.bfloat16))


@pytest.mark.parametrize("batch,tokens,heads,causal", [
    (3, 50, 2, 0), (2, 77, 8, 1), (2, 197, 12, 0), (1, 257, 16, 0), (4, 16, 2, 1), (1, 384, 1, 0), (2, 33, 2, 1),
    # T <= 128: the tcgen05 kernel (Tp = 64 / 96 / 128), several items per CTA; T > 128: the CUDA-core kernels
    (40, 50, 12, 0), (32, 77, 8, 1), (3, 128, 2, 0), (3, 128, 2, 1), (2, 96, 3, 1), (2, 65, 2, 0), (5, 1, 2, 1),
    (2, 129, 2, 1)])
def test_attention_bwd_vs_autograd(cuda_device, batch, tokens, heads, causal):
    L, lib = _lib()
    from clip_lora_match_b200 import kernels as K

    D = heads * 64
    g = torch.Generator().manual_seed(tokens * 7 + heads)
    qkv = torch.randn((batch * tokens, 3 * D), generator=g).to(cuda_device, torch.bfloat16)
    dout = torch.randn((batch * tokens, D), generator=g).to(cuda_device, torch.bfloat16)
    dqkv = torch.zeros_like(qkv)
    nb = lib.clm_attention_bwd_scratch_bytes(batch, tokens, heads)
    scratch = torch.empty(nb, dtype=torch.uint8, device=cuda_device)
    L.check(lib.clm_attention_bwd(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), scratch.data_ptr(), nb, batch,
                                  tokens, heads, causal, _st()))
    x = qkv.float().requires_grad_(True)
    q, k, v = (x.view(batch, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)[i] for i in range(3))
    s = (q @ k.transpose(-1, -2)) * 0.125
    if causal:
        s = s + torch.full((tokens, tokens), float("-inf"), device=cuda_device).triu(1)
    o = torch.softmax(s, dim=-1) @ v
    o = o.permute(0, 2, 1, 3).reshape(batch * tokens, D)
    o.backward(dout.float())
    ref = x.grad
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        got, want = dqkv[:, sl].float(), ref[:, sl]
        err, ref_n = float((got - want).norm()), float(want.norm())
        assert err <= 2e-2 * ref_n + 1e-6, f"{name}: error {err} against a gradient of norm {ref_n}"  # T = 1: dq = dk = 0
    # the forward kernel agrees with the same reference (the backward recomputes ITS softmax)
    fwd = K.attention(qkv, batch, tokens, heads, bool(causal)).float()
    assert float((fwd - o.detach()).norm() / o.detach().norm()) < 2e-2
    # deterministic
    dqkv2 = torch.zeros_like(qkv)
    L.check(lib.clm_attention_bwd(qkv.data_ptr(), dout.data_ptr(), dqkv2.data_ptr(), scratch.data_ptr(), nb, batch,
                                  tokens, heads, causal, _st()))
    assert torch.equal(dqkv, dqkv2)


def test_clip_loss_vs_reference_golden(cuda_device):
    """loss and d loss / d features against the outputs of the reference's own compute_clip_contrastive_loss
    (tests/golden/train_golden.npz, oracle/make_golden.py)."""
    L, lib = _lib()
    g = np.load(os.path.join(GOLD, "train_golden.npz"))
    for ci in range(int(g["n_cases"])):
        n, d = (int(v) for v in g[f"c{ci}_shape"])
        fi = torch.tensor(g[f"c{ci}_fi"]).to(cuda_device)
        ft = torch.tensor(g[f"c{ci}_ft"]).to(cuda_device)
        nb = lib.clm_clip_loss_workspace_bytes(n, d)
        ws = torch.empty(nb, dtype=torch.uint8, device=cuda_device)
        loss = torch.zeros(1, device=cuda_device)
        dfi, dft = torch.zeros_like(fi), torch.zeros_like(ft)
        dfi_b = torch.zeros((n, d), dtype=torch.bfloat16, device=cuda_device)
        dft_b = torch.zeros((n, d), dtype=torch.bfloat16, device=cuda_device)
        for scale in (1.0, 0.25):
            L.check(lib.clm_clip_loss(fi.data_ptr(), ft.data_ptr(), n, d, float(g[f"c{ci}_temp"]), scale,
                                      loss.data_ptr(), dfi.data_ptr(), dft.data_ptr(), dfi_b.data_ptr(),
                                      dft_b.data_ptr(), ws.data_ptr(), nb, _st()))
            want = float(g[f"c{ci}_loss"]) * scale
            assert abs(loss.item() - want) <= 2e-5 * max(1.0, abs(want)), (ci, loss.item(), want)
            for got, ref in ((dfi, g[f"c{ci}_dfi"]), (dft, g[f"c{ci}_dft"])):
                ref = torch.tensor(ref).to(cuda_device) * scale
                assert float((got - ref).norm() / ref.norm()) < 1e-4
            assert torch.equal(dfi_b, dfi.to(torch.bfloat16))
        # forward only (validation): no gradient buffers
        L.check(lib.clm_clip_loss(fi.data_ptr(), ft.data_ptr(), n, d, float(g[f"c{ci}_temp"]), 1.0, loss.data_ptr(),
                                  None, None, None, None, ws.data_ptr(), nb, _st()))
        assert abs(loss.item() - float(g[f"c{ci}_loss"])) <= 2e-5 * max(1.0, float(g[f"c{ci}_loss"]))


def test_clip_loss_large_batch(cuda_device):
    L, lib = _lib()
    n, d = 1024, 512
    g = torch.Generator().manual_seed(5)
    fi = torch.randn((n, d), generator=g)
    ft = 0.2 * fi + torch.randn((n, d), generator=g)
    a, b = fi.clone().requires_grad_(True), ft.clone().requires_grad_(True)
    ref = T.contrastive_loss(a, b, 0.07)
    ref.backward()
    fi_d, ft_d = fi.to(cuda_device), ft.to(cuda_device)
    nb = lib.clm_clip_loss_workspace_bytes(n, d)
    ws = torch.empty(nb, dtype=torch.uint8, device=cuda_device)
    loss = torch.zeros(1, device=cuda_device)
    dfi, dft = torch.zeros_like(fi_d), torch.zeros_like(ft_d)
    L.check(lib.clm_clip_loss(fi_d.data_ptr(), ft_d.data_ptr(), n, d, 0.07, 1.0, loss.data_ptr(), dfi.data_ptr(),
                              dft.data_ptr(), None, None, ws.data_ptr(), nb, _st()))
    assert abs(loss.item() - ref.item()) < 1e-4
    assert float((dfi.cpu() - a.grad).norm() / a.grad.norm()) < 1e-3
    assert float((dft.cpu() - b.grad).norm() / b.grad.norm()) < 1e-3


@pytest.mark.parametrize("max_norm", [0.0, 1.0])
def test_adamw_step_vs_torch(cuda_device, max_norm):
    L, lib = _lib()
    n = 10000
    g = torch.Generator().manual_seed(9)
    p0 = torch.randn(n, generator=g)
    mult = torch.ones(n)
    mult[::7] = 0.0     # structural zeros
    mult[1::5] = 2.0    # LoRA scaling on B entries
    p0[mult == 0] = 0.0
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p], lr=1e-3, weight_decay=0.01)
    pd = p0.clone().to(cuda_device)
    md = mult.to(cuda_device)
    m = torch.zeros(n, device=cuda_device)
    v = torch.zeros(n, device=cuda_device)
    hyper = torch.zeros(4, device=cuda_device)
    sumsq = torch.zeros(1185, device=cuda_device)
    for step in range(1, 4):
        graw = torch.randn(n, generator=g) * 3.0
        p.grad = graw * mult
        if max_norm > 0:
            norm = torch.nn.utils.clip_grad_norm_([p], max_norm)
        else:
            norm = p.grad.norm()
        opt.step()
        hyper.copy_(torch.tensor([1e-3, 1 / (1 - 0.9 ** step), 1 / (1 - 0.999 ** step) ** 0.5, 0.0]))
        gd = graw.to(cuda_device)
        L.check(lib.clm_adamw_step(pd.data_ptr(), gd.data_ptr(), md.data_ptr(), m.data_ptr(), v.data_ptr(), n,
                                   hyper.data_ptr(), sumsq.data_ptr(), max_norm, 0.9, 0.999, 1e-8, 0.01, _st()))
        assert abs(sumsq[0].item() ** 0.5 - float(norm)) < 1e-3 * float(norm)
        assert torch.allclose(pd.cpu(), p.detach(), rtol=1e-5, atol=1e-6), (pd.cpu() - p.detach()).abs().max()
        assert not pd[md == 0].any()


# ------------------------------------------------------------------------------------------------------------
# the step against the autograd oracle
# ------------------------------------------------------------------------------------------------------------
def _pair(arch_name, r, alpha, targets, device, seed=0):
    from clip_lora_match_b200.models import clip_model as CM
    from clip_lora_match_b200.models.lora_adapter import LoraAdapter, LoraConfig

    model = O.build_model(arch_name, seed=seed)
    weights = O.synthetic_lora(model, r, alpha, targets, seed=1)
    arch = (CM.arch_from_hf_config(O.hf_config(arch_name), arch_name) if arch_name == "tiny-test"
            else CM.arch_from_name(arch_name))
    lora = LoraAdapter(LoraConfig(r=r, lora_alpha=alpha, lora_dropout=0.0, target_modules=list(targets)),
                       {k: (a.clone(), b.clone()) for k, (a, b) in weights.items()})
    gpu = CM.B200ClipModel(arch, O.base_state_dict(model), lora=lora, device=device)
    return model, gpu


STEP_CASES = [
    ("tiny-test", 6, 8, 16, ("q_proj", "k_proj", "v_proj", "out_proj")),   # the shipped YAML's targets
    ("tiny-test", 5, 8, 16, ("q_proj", "v_proj")),
    ("tiny-test", 4, 8, 16, ("fc1", "fc2", "out_proj")),
    ("tiny-test", 4, 32, 32, ("q_proj", "k_proj", "v_proj")),               # 96 -> 128 LoRA columns
    ("openai/clip-vit-base-patch32", 8, 8, 16, ("q_proj", "k_proj", "v_proj", "out_proj")),  # config/lora_config.yaml
]


@pytest.mark.parametrize("arch,batch,r,alpha,targets", STEP_CASES)
def test_gradients_vs_autograd_oracle(cuda_device, arch, batch, r, alpha, targets):
    from clip_lora_match_b200.models.lora_trainer import LoraTrainer

    torch.set_num_threads(os.cpu_count() or 8)
    model, gpu = _pair(arch, r, alpha, targets, cuda_device)
    pv = O.synth_images(batch, seed=2)
    ids, mask = O.synth_captions(batch, seed=3)
    oracle = T.TrainOracle(model, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0, temperature=0.07)
    ref_loss = oracle.forward_backward(pv, ids, mask)
    ref_grads = oracle.gradients()
    tr = LoraTrainer(gpu, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0, temperature=0.07, use_graph=False)
    assert tr.num_trainable_parameters() == sum(p.numel() for p in oracle.params)
    loss = tr.forward_backward(pv.to(cuda_device), ids.to(cuda_device), mask.to(cuda_device)).item()
    print(f"[train] {arch} loss gpu {loss:.5f} oracle {ref_loss:.5f}")
    assert abs(loss - ref_loss) < 2e-2, (loss, ref_loss)
    # features of the training forward = the oracle's get_*_features
    with torch.no_grad():
        fi, ft = oracle.features(pv, ids, mask)
    gi, gt = tr.features()
    assert O.parity_metrics(gi.cpu(), fi)["cos_min"] >= 0.999
    assert O.parity_metrics(gt.cpu(), ft)["cos_min"] >= 0.999
    grads = tr.gradients()
    assert set(grads) == set(ref_grads)
    flat_g, flat_r, worst = [], [], (1.0, "")
    total = sum(float(a.norm() ** 2 + b.norm() ** 2) for a, b in ref_grads.values()) ** 0.5
    for path, (ra, rb) in ref_grads.items():
        ga, gb = grads[path]
        for tag, got, ref in (("A", ga, ra), ("B", gb, rb)):
            flat_g.append(got.flatten()); flat_r.append(ref.flatten())
            if float(ref.norm()) > 1e-3 * total:  # a factor that matters for the update
                c = _cos(got, ref)
                if c < worst[0]:
                    worst = (c, f"{path}.{tag}")
                assert abs(float(got.norm()) / float(ref.norm()) - 1.0) < 0.1, f"{path}.{tag} norm"
    all_cos = _cos(torch.cat(flat_g), torch.cat(flat_r))
    print(f"[train] {arch} gradient cosine overall {all_cos:.5f}, worst factor {worst}")
    assert all_cos >= 0.99
    assert worst[0] >= 0.98, worst
    # clip + AdamW: same update as torch's on the oracle's gradients
    before = {k: (a.clone(), b.clone()) for k, (a, b) in tr.export_adapter().weights.items()}
    ref_norm = oracle.optimizer_step(lr=1e-3)
    tr.optimizer_step(lr=1e-3)
    assert abs(tr.sumsq[0].item() ** 0.5 - ref_norm) < 0.03 * ref_norm
    after = tr.export_adapter().weights
    ref_after = O.get_lora_weights(model)
    du_g = torch.cat([torch.cat([(after[k][0] - before[k][0]).flatten(), (after[k][1] - before[k][1]).flatten()]) for k in after])
    du_r = torch.cat([torch.cat([(ref_after[k][0] - before[k][0]).flatten(), (ref_after[k][1] - before[k][1]).flatten()]) for k in after])
    # the first AdamW update is lr * sign-like(g): compare where the oracle's gradient is not tiny
    print(f"[train] {arch} update cosine {_cos(du_g, du_r):.4f}")
    assert _cos(du_g, du_r) >= 0.9
    assert not tr.grad.any()  # zero_grad


def test_training_trajectory_graph_and_export(cuda_device, tmp_path):
    """Six optimizer steps on a fixed batch: the loss falls as the oracle's does, the CUDA-graph replay path gives
    the eager path's numbers bit for bit, and the exported adapter (PEFT layout, train_lora.py:243-247) reloads
    into the inference model and reproduces the trainer's features."""
    from clip_lora_match_b200.models import clip_model as CM
    from clip_lora_match_b200.models.lora_adapter import load_lora_adapter, save_lora_adapter
    from clip_lora_match_b200.models.lora_trainer import LoraTrainer

    torch.set_num_threads(os.cpu_count() or 8)
    targets = ("q_proj", "k_proj", "v_proj", "out_proj")
    batch = 8
    pv = O.synth_images(batch, seed=2)
    ids, mask = O.synth_captions(batch, seed=3)
    model, gpu = _pair("tiny-test", 8, 16, targets, cuda_device)
    oracle = T.TrainOracle(model, lr=2e-3, weight_decay=0.01, max_grad_norm=1.0, temperature=0.07)
    ref_losses = [oracle.step(pv, ids, mask, lr=2e-3) for _ in range(6)]
    runs = {}
    for use_graph in (False, True):
        _, gpu = _pair("tiny-test", 8, 16, targets, cuda_device)
        tr = LoraTrainer(gpu, lr=2e-3, weight_decay=0.01, max_grad_norm=1.0, temperature=0.07, use_graph=use_graph,
                         deterministic=True)
        losses = [tr.step(pv.to(cuda_device), ids.to(cuda_device), mask.to(cuda_device), lr=2e-3).item() for _ in range(6)]
        runs[use_graph] = (losses, tr.theta.clone(), tr)
    print(f"[train] oracle {ref_losses}\n[train] eager  {runs[False][0]}\n[train] graph  {runs[True][0]}")
    assert runs[True][0] == runs[False][0]
    assert torch.equal(runs[True][1], runs[False][1])
    assert runs[True][2]._graph_state == 2
    for got, ref in zip(runs[False][0], ref_losses):
        assert abs(got - ref) < 0.05 + 0.05 * abs(ref)
    assert runs[False][0][-1] < runs[False][0][0] - 0.1
    # export -> PEFT directory -> inference model
    tr = runs[True][2]
    val = tr.eval_loss(pv.to(cuda_device), ids.to(cuda_device), mask.to(cuda_device)).item()
    assert val < runs[True][0][0]
    fi = tr.features()[0].clone()
    out_dir = save_lora_adapter(tr.export_adapter(), tmp_path / "epoch_1")
    assert (out_dir / "adapter_config.json").exists() and (out_dir / "adapter_model.safetensors").exists()
    gpu.set_lora(load_lora_adapter(out_dir))
    raw = gpu.encode_images(pv.to(cuda_device), normalize=False)
    assert O.parity_metrics(raw.cpu(), fi.cpu())["rel_l2_max"] < 2e-2


def test_gradient_accumulation_matches_one_big_loss_scale(cuda_device):
    """Two micro-batches with grad_accum_steps = 2 accumulate loss / 2 gradients (train_lora.py:186-190)."""
    from clip_lora_match_b200.models.lora_trainer import LoraTrainer

    targets = ("q_proj", "v_proj")
    model, gpu = _pair("tiny-test", 8, 16, targets, cuda_device)
    oracle = T.TrainOracle(model, lr=1e-3, grad_accum_steps=2)
    tr = LoraTrainer(gpu, lr=1e-3, grad_accum_steps=2, use_graph=False)
    for seed in (2, 12):
        pv = O.synth_images(4, seed=seed)
        ids, mask = O.synth_captions(4, seed=seed + 1)
        ref = oracle.forward_backward(pv, ids, mask)
        got = tr.forward_backward(pv.to(cuda_device), ids.to(cuda_device), mask.to(cuda_device)).item()
        assert abs(got - ref) < 1e-2
    ref_grads, grads = oracle.gradients(), tr.gradients()
    fg = torch.cat([torch.cat([grads[k][0].flatten(), grads[k][1].flatten()]) for k in ref_grads])
    fr = torch.cat([torch.cat([ref_grads[k][0].flatten(), ref_grads[k][1].flatten()]) for k in ref_grads])
    assert _cos(fg, fr) >= 0.99
    assert abs(float(fg.norm() / fr.norm()) - 1.0) < 0.05


def test_train_script_mirror_end_to_end(cuda_device, tmp_path):
    """scripts/train_lora.train() on synthetic loaders: two epochs with gradient accumulation off and on, per-epoch
    PEFT checkpoints (reference :243-247) that load_clip_model(use_lora=True) accepts, and
    compute_clip_contrastive_loss equal to the oracle's restatement of the reference's."""
    import yaml

    from clip_lora_match_b200.models import clip_model as CM
    from clip_lora_match_b200.scripts import train_lora as TL

    arch = CM.arch_from_hf_config(O.hf_config("tiny-test"), "tiny-test")
    sd = O.base_state_dict(O.build_model("tiny-test", seed=0))

    def batches(n, seed):
        out = []
        for k in range(n):
            ids, mask = O.synth_captions(8, seed=seed + 2 * k + 1)
            out.append({"pixel_values": O.synth_images(8, seed=seed + 2 * k), "input_ids": ids, "attention_mask": mask})
        return out

    for accum in (1, 2):
        cfg = {"model": {"base_model_name": "tiny-test", "target_modules": ["q_proj", "k_proj", "v_proj", "out_proj"]},
               "lora": {"r": 8, "alpha": 16, "dropout": 0.1},
               "training": {"seed": 1, "batch_size": 8, "learning_rate": 2e-3, "num_epochs": 2, "logging_steps": 2,
                            "gradient_accumulation_steps": accum, "warmup_ratio": 0.25,
                            "output_dir": str(tmp_path / f"out{accum}")}}
        ypath = tmp_path / f"lora{accum}.yaml"
        ypath.write_text(yaml.safe_dump(cfg))
        model = CM.B200ClipModel(arch, sd, lora=None, device=cuda_device)
        train, val = batches(4, 100), batches(1, 900)
        pv, ids, mask = (val[0][k].to(cuda_device) for k in ("pixel_values", "input_ids", "attention_mask"))
        trainer = TL.train(ypath, loaders=(train, val), model=model)
        assert trainer.opt_step == 2 * 4 // accum
        for e in (1, 2):
            d = tmp_path / f"out{accum}" / f"epoch_{e}"
            assert (d / "adapter_config.json").exists() and (d / "adapter_model.safetensors").exists()
        # B is zero at attach time: any non-zero B proves the optimizer moved the factors
        w = trainer.export_adapter().weights
        assert all(float(b.abs().max()) > 0 for _, b in w.values())
        # the trained model (sync_model) reproduces the trainer's own features
        trainer.eval_loss(pv, ids, mask)
        fi, ft = (f.clone() for f in trainer.features())
        assert O.parity_metrics(model.encode_images(pv, normalize=False).cpu(), fi.cpu())["rel_l2_max"] < 2e-2
        assert O.parity_metrics(model.encode_texts(ids, normalize=False).cpu(), ft.cpu())["rel_l2_max"] < 2e-2
        got = TL.compute_clip_contrastive_loss(fi, ft, 0.07).item()
        assert abs(got - float(T.contrastive_loss(fi.cpu(), ft.cpu(), 0.07))) < 1e-4


def test_split_k_weight_gradients(cuda_device):
    """Deep-K weight-gradient GEMMs with CLM_EPI_SPLIT_K (several work units per tile adding through the L2) give the
    one-unit-per-tile result up to fp32 rounding, at a batch whose token count (12,800 / 19,712 rows) triggers it."""
    from clip_lora_match_b200 import _lib as L
    from clip_lora_match_b200 import kernels as K

    g = torch.Generator().manual_seed(11)
    rows = 19712
    for feats in (512, 2304):
        dyT = torch.randn((feats, rows), generator=g).to(cuda_device, torch.bfloat16)
        tT = torch.randn((64, rows), generator=g).to(cuda_device, torch.bfloat16)
        bias = torch.randn(64, generator=g).to(cuda_device)
        base = torch.randn((feats, 64), generator=g).to(cuda_device)
        ref = base.double() + dyT.double() @ tT.double().t() + bias.double()
        outs = []
        for flag in (L.EPI_NONE, L.EPI_SPLIT_K):
            out = base.clone()
            K.gemm_epi(dyT, tT, bias=bias, residual=out, out=out, act=flag)
            outs.append(out)
            rel = float((out.double() - ref).norm() / ref.norm())
            assert rel < 1e-4, (feats, flag, rel)
        diff = float((outs[0] - outs[1]).abs().max())
        assert diff < 2e-3 * float(ref.abs().max()), (feats, diff)
    # the flag is ignored where it does not apply (bf16 store): same bits as without it
    a = torch.randn((300, 4096), generator=g).to(cuda_device, torch.bfloat16)
    w = torch.randn((64, 4096), generator=g).to(cuda_device, torch.bfloat16)
    assert torch.equal(K.gemm_epi(a, w, act=L.EPI_SPLIT_K), K.gemm_epi(a, w))


def test_lora_wgrad_small_vs_torch(cuda_device):
    L, lib = _lib()
    g = torch.Generator().manual_seed(13)
    for rows, n_out, n_in, cols in [(400, 2304, 768, 64), (616, 512, 512, 64), (77, 256, 128, 128), (33, 130, 66, 64)]:
        dy = torch.randn((rows, n_out), generator=g).to(cuda_device, torch.bfloat16)
        t = torch.randn((rows, cols), generator=g).to(cuda_device, torch.bfloat16)
        x = torch.randn((rows, n_in), generator=g).to(cuda_device, torch.bfloat16)
        u = torch.randn((rows, cols), generator=g).to(cuda_device, torch.bfloat16)
        gb0 = torch.randn((n_out, cols), generator=g).to(cuda_device)
        ga0 = torch.randn((n_in, cols), generator=g).to(cuda_device)
        for det in (1, 0):  # one CTA per tile / up to four CTAs per tile adding with atomics
            gb, ga = gb0.clone(), ga0.clone()
            L.check(lib.clm_lora_wgrad_small(dy.data_ptr(), n_out, n_out, t.data_ptr(), cols, x.data_ptr(), n_in, n_in,
                                             u.data_ptr(), cols, cols, rows, gb.data_ptr(), ga.data_ptr(), det, _st()))
            assert torch.allclose(gb, gb0 + dy.float().t() @ t.float(), rtol=1e-4, atol=1e-3)
            assert torch.allclose(ga, ga0 + x.float().t() @ u.float(), rtol=1e-4, atol=1e-3)


def test_small_and_tensor_core_weight_gradients_agree(cuda_device):
    """The same micro-batch through the one-launch CUDA-core weight-gradient kernel (few rows) and through the
    transposes + tcgen05 GEMMs (forced): the same gradients up to bf16-product / fp32-order rounding."""
    from clip_lora_match_b200.models.lora_trainer import LoraTrainer

    targets = ("q_proj", "k_proj", "v_proj", "out_proj")
    pv = O.synth_images(6, seed=2)
    ids, mask = O.synth_captions(6, seed=3)
    flats = []
    for small_rows in (LoraTrainer.SMALL_WGRAD_ROWS, 0):
        _, gpu = _pair("tiny-test", 8, 16, targets, cuda_device)
        tr = LoraTrainer(gpu, use_graph=False, deterministic=True)
        tr.SMALL_WGRAD_ROWS = small_rows
        tr.forward_backward(pv.to(cuda_device), ids.to(cuda_device), mask.to(cuda_device))
        flats.append((tr.grad * tr.mult).clone())
    assert _cos(flats[0], flats[1]) > 0.999999
    assert float((flats[0] - flats[1]).abs().max()) <= 1e-4 * float(flats[1].abs().max()) + 1e-7
