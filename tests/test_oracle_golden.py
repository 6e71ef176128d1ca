"""CPU tests: the oracle (oracle/clip_oracle.py) against golden vectors produced by the
reference's own code (oracle/make_golden.py, run in the build container where /root/reference
is mounted).  These pin the oracle before any GPU parity claim is made against it.
"""
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sg():
    return np.load(os.path.join(GOLD, "search_golden.npz"))


@pytest.fixture(scope="module")
def eg():
    return np.load(os.path.join(GOLD, "encoder_golden.npz"))


def test_fixture_rows_are_unit_norm(sg):
    e = torch.from_numpy(sg["fixture_embeddings"])
    assert e.shape == (6, 512)
    assert torch.allclose(e.norm(dim=-1), torch.ones(6), atol=1e-5)


@pytest.mark.parametrize("k", [1, 3, 5, 10])
def test_search_matches_reference_on_shipped_fixture(sg, k):
    """reference TextSearchIndex.search_with_embedding over data/index/custom_items_index.pt."""
    e = torch.from_numpy(sg["fixture_embeddings"])
    q = torch.from_numpy(sg["fixture_queries"])
    s, i = O.search_topk(e, q, k)
    assert i.shape == sg[f"fixture_ids_k{k}"].shape  # k is clamped to N=6 (search.py:98)
    assert np.array_equal(i.numpy(), sg[f"fixture_ids_k{k}"])
    assert np.allclose(s.numpy(), sg[f"fixture_scores_k{k}"], atol=1e-6)


def test_search_matches_reference_on_synthetic_index(sg):
    n, d, nq = int(sg["synth_n"]), int(sg["synth_d"]), int(sg["synth_nq"])
    emb = O.synth_unit_rows(n, d, 4) * 1.7
    qs = torch.randn((nq, d), generator=torch.Generator().manual_seed(5))
    s, i = O.search_topk(emb, qs, 10)
    assert np.array_equal(i.numpy(), sg["synth_ids_k10"])
    assert np.allclose(s.numpy(), sg["synth_scores_k10"], atol=1e-6)
    # reference similarity.top_k_similar / cosine_similarity
    s5, i5 = O.search_topk(emb, qs[:8], 5)
    assert np.array_equal(i5.numpy(), sg["sim_indices_k5"])
    assert np.allclose(s5.numpy(), sg["sim_values_k5"], atol=1e-6)
    cos = (O.normalize_rows(qs[:1]) @ O.normalize_rows(emb).T)[0, :64]
    assert np.allclose(cos.numpy(), sg["sim_cosine_q0"], atol=1e-6)


def _rebuild_case(eg, ci):
    pre = f"case{ci}_"
    arch = str(eg[pre + "arch"])
    targets = str(eg[pre + "targets"]).split(",")
    r, alpha = int(eg[pre + "r"]), int(eg[pre + "alpha"])
    model = O.build_model(arch, seed=0)
    weights = O.synthetic_lora(model, r, alpha, targets, seed=1, b_std=0.02)
    return pre, arch, model, weights


def _golden_pixels(seeds):
    from transformers import CLIPImageProcessor

    ip = CLIPImageProcessor()
    imgs = [np.random.RandomState(int(s)).randint(0, 256, size=(224, 224, 3), dtype=np.uint8) for s in seeds]
    return ip(images=imgs, return_tensors="pt")["pixel_values"]


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_encoder_oracle_matches_reference_code(eg, ci):
    """The oracle's batched encode == the reference's encode_image/encode_text (one item at a
    time, unpadded captions) with LoRA attached by the reference's attach_lora_to_clip."""
    torch.set_num_threads(max(1, (os.cpu_count() or 2)))
    pre, arch, model, weights = _rebuild_case(eg, ci)
    assert len(weights) == int(eg[pre + "n_wrapped"])
    r = int(eg[pre + "r"])
    trainable = sum(a.numel() + b.numel() for a, b in weights.values())
    assert trainable == int(eg[pre + "trainable"])
    pv = _golden_pixels(eg[pre + "image_seeds"])
    img = O.encode_images(model, pv)
    assert np.allclose(img.numpy(), eg[pre + "image_emb"], atol=2e-5)
    ids = torch.from_numpy(eg[pre + "input_ids"])
    n_txt = eg[pre + "text_emb"].shape[0]
    # padded + masked batch == the reference's unpadded single captions (causal tower)
    mask = (torch.arange(77).unsqueeze(0) <= (ids == O.EOS_ID).int().argmax(dim=-1, keepdim=True)).long()
    txt = O.encode_texts(model, ids[:n_txt], mask[:n_txt])
    assert np.allclose(txt.numpy(), eg[pre + "text_emb"], atol=2e-5)
    # and without a mask at all (what the CUDA path does: causal mask only)
    txt2 = O.encode_texts(model, ids[:n_txt], None)
    assert np.allclose(txt2.numpy(), eg[pre + "text_emb"], atol=2e-5)


def test_lora_is_not_a_noop_and_merges():
    """W' = W + (alpha/r) B A reproduces the unmerged forward (Appendix B cross-check)."""
    model = O.build_model("tiny-test", seed=0)
    pv = O.synth_images(2, seed=2)
    base = O.encode_images(model, pv)
    O.synthetic_lora(model, 8, 16, ["q_proj", "v_proj"], seed=1, b_std=0.05)
    with_lora = O.encode_images(model, pv)
    assert (base - with_lora).abs().max() > 1e-4
    merged = O.build_model("tiny-test", seed=0)
    sd = merged.state_dict()
    for path, (a, b) in O.get_lora_weights(model).items():
        sd[path + ".weight"] += (16 / 8) * (b @ a)
    merged.load_state_dict(sd)
    assert torch.allclose(O.encode_images(merged, pv), with_lora, atol=1e-5)


def test_synth_captions_shape_and_framing():
    ids, mask = O.synth_captions(64, seed=3)
    assert ids.shape == (64, 77) and mask.shape == (64, 77)
    assert (ids[:, 0] == O.BOS_ID).all()
    lengths = mask.sum(dim=1)
    assert lengths.min() >= 3 and lengths.max() <= 77
    first_eos = (ids == O.EOS_ID).int().argmax(dim=-1)
    assert torch.equal(first_eos, lengths - 1)


def test_fusion_oracle_matches_reference_seeker_service():
    """oracle.fuse_query == the reference's SeekerService._build_query_embedding (bit for bit: same torch ops)."""
    g = np.load(os.path.join(GOLD, "fusion_golden.npz"))
    txt, img = torch.from_numpy(g["text"]), torch.from_numpy(g["image"])
    for wi, (wt, wim) in enumerate(g["weights"].tolist()):
        got = torch.stack([O.fuse_query(txt[i], img[i], wt, wim) for i in range(txt.shape[0])])
        assert torch.equal(got, torch.from_numpy(g["both"][wi]))
        assert torch.allclose(O.fuse_query(txt, img, wt, wim), torch.from_numpy(g["both"][wi]), atol=1e-7)
    assert torch.equal(torch.stack([O.fuse_query(t, None) for t in txt]), torch.from_numpy(g["only_text"]))
    assert torch.equal(torch.stack([O.fuse_query(None, t) for t in img]), torch.from_numpy(g["only_image"]))
    with pytest.raises(ValueError, match=str(g["error"])):
        O.fuse_query(None, None)


# ------------------------------------------------------------------------------------------
# training step: the oracle's loss and schedule against the reference's own functions
# ------------------------------------------------------------------------------------------
def test_train_oracle_loss_and_schedule_match_reference():
    import torch

    from oracle import train_oracle as T

    g = np.load(os.path.join(GOLD, "train_golden.npz"))
    for ci in range(int(g["n_cases"])):
        fi = torch.tensor(g[f"c{ci}_fi"], requires_grad=True)
        ft = torch.tensor(g[f"c{ci}_ft"], requires_grad=True)
        loss = T.contrastive_loss(fi, ft, float(g[f"c{ci}_temp"]))
        loss.backward()
        assert abs(loss.item() - float(g[f"c{ci}_loss"])) <= 1e-6 * max(1.0, abs(float(g[f"c{ci}_loss"])))
        assert np.allclose(fi.grad.numpy(), g[f"c{ci}_dfi"], rtol=1e-5, atol=1e-8)
        assert np.allclose(ft.grad.numpy(), g[f"c{ci}_dft"], rtol=1e-5, atol=1e-8)
    for row in g["sched"]:
        total, warm = int(row[0]), int(row[1])
        for s in range(total + 2):
            assert T.lr_lambda(s, total, warm) == row[2 + s]
