"""CPU test of the N>1 host path: world_size-2 gloo process group, row-sharded index, local
top-k per rank (the oracle stands in for the CUDA scan here), the ONE exchange step
(gather_shard_topk = one all_gather_into_tensor of the packed Q*k pairs) and the merge == unsharded oracle search.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clip_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, nq, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clip_lora_match_b200.src.embedding.search import gather_shard_topk, shard_bounds, unpack_gathered_topk

        index = O.synth_unit_rows(n, d, 4)
        queries = O.synth_unit_rows(nq, d, 5)
        lo, hi = shard_bounds(n, rank, world)
        k_local = min(k, hi - lo)
        if k_local > 0:
            s, i = O.search_topk(index[lo:hi], queries, k_local)
            i = i + lo
        else:
            s = i = None
        from clip_lora_match_b200.src.embedding.search import gather_overflow_counts

        # the per-rank "queries to redo exactly" counts reach every rank identically (rank 1 pretends it has two)
        flags = torch.zeros(nq, dtype=torch.int32)
        if rank == 1:
            flags[:2] = 1
        assert gather_overflow_counts(flags if rank else None, torch.device("cpu")) == [0, 2][:world]
        buf = gather_shard_topk(s, i, nq, k, torch.device("cpu"))  # ONE all_gather_into_tensor of packed chunks
        assert buf.dtype == torch.uint8 and buf.numel() == world * (((nq * k + 1) // 2 * 2) * 12)
        gs, gi = unpack_gathered_topk(buf, world, nq, k)
        assert gs.shape == (nq, world, k) and gi.shape == (nq, world, k)
        # every rank merges the same candidates; checker merge = torch.topk over the gathered lists
        flat_s, flat_i = gs.reshape(nq, -1), gi.reshape(nq, -1)
        ms, pos = torch.topk(flat_s, min(k, n), dim=-1)
        mi = torch.gather(flat_i, 1, pos)
        ref_s, ref_i = O.search_topk(index, queries, k)
        ok = torch.equal(mi, ref_i) and torch.allclose(ms, ref_s, atol=1e-6)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _run(n, d, nq, k, world=2):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, d, nq, k, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_sharded_search_two_ranks_matches_unsharded():
    _run(n=4001, d=64, nq=9, k=10)


def test_sharded_search_shard_smaller_than_k():
    # 5 rows over 2 ranks: shards of 2 and 3 rows, k=4 > shard size -> padded payload
    _run(n=5, d=32, nq=3, k=4)


def _gather_worker(rank, world, port, total, d, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clip_lora_match_b200.src.embedding.search import allgather_rows, shard_bounds

        full = O.synth_unit_rows(total, d, 11)
        lo, hi = shard_bounds(total, rank, world)
        got = allgather_rows(full[lo:hi].clone(), total)
        bad = False
        try:
            allgather_rows(full[lo:hi + 1].clone(), total) if hi < total else allgather_rows(full[lo:hi - 1].clone(), total)
        except ValueError:
            bad = True
        ret[rank] = bool(torch.equal(got, full)) and bad
    finally:
        dist.destroy_process_group()


def test_query_rows_are_split_over_ranks_and_gathered_back():
    """Seeker path, query side: each rank encodes its block of the query batch; allgather_rows rebuilds the
    [Q, d] matrix on every rank (uneven split: 7 rows over 2 ranks), and rejects a block of the wrong size."""
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gather_worker, args=(2, port, 7, 16, ret), nprocs=2, join=True)
    assert all(ret.get(r) for r in range(2)), dict(ret)


# ------------------------------------------------------------------------------------------------------------
# data-parallel LoRA training: the one exchange step (models/lora_trainer.allreduce_gradients)
# ------------------------------------------------------------------------------------------------------------
def _train_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clip_lora_match_b200.models.lora_trainer import allreduce_gradients
        from oracle import train_oracle as T

        torch.set_num_threads(2)
        targets = ("q_proj", "v_proj")

        def flat(oracle):
            return torch.cat([torch.cat([a.flatten(), b.flatten()]) for a, b in oracle.gradients().values()])

        # every rank: its own micro-batch, loss scaled by 1 / world (the oracle stands in for the CUDA step here)
        model = O.build_model("tiny-test", seed=0)
        O.synthetic_lora(model, 8, 16, targets, seed=1)
        oracle = T.TrainOracle(model, grad_accum_steps=world)
        pv = O.synth_images(3, seed=20 + rank)
        ids, mask = O.synth_captions(3, seed=30 + rank)
        loss = torch.tensor([oracle.forward_backward(pv, ids, mask)])
        grad = flat(oracle)
        allreduce_gradients(grad, loss)  # ONE all-reduce of the flat gradient (+ 4 bytes of loss)
        # checker: one process accumulating the same micro-batches (train_lora.py:186-190)
        model2 = O.build_model("tiny-test", seed=0)
        O.synthetic_lora(model2, 8, 16, targets, seed=1)
        acc = T.TrainOracle(model2, grad_accum_steps=world)
        total = 0.0
        for r in range(world):
            total += acc.forward_backward(O.synth_images(3, seed=20 + r), *O.synth_captions(3, seed=30 + r))
        ref = flat(acc)
        ret[rank] = bool(torch.allclose(grad, ref, rtol=1e-4, atol=1e-7) and abs(loss.item() - total) < 1e-5)
    finally:
        dist.destroy_process_group()


def test_data_parallel_training_exchange_equals_gradient_accumulation():
    world, port = 2, _free_port()
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_train_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_allreduce_gradients_is_a_noop_without_a_process_group():
    from clip_lora_match_b200.models.lora_trainer import allreduce_gradients

    g = torch.arange(8, dtype=torch.float32)
    allreduce_gradients(g, torch.ones(1))
    assert torch.equal(g, torch.arange(8, dtype=torch.float32))
