"""Multi-GPU check of the row-sharded search (run under torchrun, one rank per GPU): every rank builds the same
random unit-norm index, keeps its shard, and search_batch's merged top-k must equal torch.topk of the fp32 scores
of the WHOLE index (computed by each rank on its own GPU as the checker) up to 1e-4 ties; also prints ms per batch.
Usage: torchrun --nproc-per-node N tools/dist_search_check.py [--rows 2000000] [--queries 1024] [--topk 10 50]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", dest="n", type=int, default=2_000_000); ap.add_argument("--queries", dest="q", type=int, default=1024)
    ap.add_argument("--dim", dest="d", type=int, default=768); ap.add_argument("--topk", dest="k", type=int, nargs="+", default=[10, 50])
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda")
    dist.init_process_group("nccl")
    from clip_lora_match_b200.src.embedding.search import TextSearchIndex, shard_bounds
    g = torch.Generator(device=dev).manual_seed(4)
    e = torch.randn((a.n, a.d), generator=g, device=dev); e = e / e.norm(dim=-1, keepdim=True)
    q = torch.randn((a.q, a.d), generator=g, device=dev); q = q / q.norm(dim=-1, keepdim=True)
    lo, hi = shard_bounds(a.n, rank, world)
    idx = TextSearchIndex(embeddings=e[lo:hi].clone(), device=dev, distributed=True, row_offset=lo, total_rows=a.n, verbose=False)
    sims = q @ e.T                                  # checker: fp32 scores of the whole index
    ok_all = True
    for k in a.k:
        s, i = idx.search_batch(q, top_k=k)
        ref_s, ref_i = torch.topk(sims, k, dim=-1)
        got_s = torch.gather(sims, 1, i)            # the fp32 checker score of every returned id
        ok = bool(torch.allclose(s, ref_s, atol=3e-6)) and bool((got_s >= ref_s[:, -1:] - 1e-4).all()) \
            and all(len(set(r)) == k for r in i[:8].tolist())
        for _ in range(2): idx.search_batch(q, top_k=k)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): idx.search_batch(q, top_k=k)
        e1.record(); torch.cuda.synchronize()
        ok_all &= ok
        if rank == 0:
            print(json.dumps({"world": world, "n": a.n, "q": a.q, "k": k, "ids_and_scores_match": ok,
                              "ms_per_batch": round(e0.elapsed_time(e1) / 5, 3), "stats": idx.last_search_stats}), flush=True)
    flag = torch.tensor([int(ok_all)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)

if __name__ == "__main__":
    main()
