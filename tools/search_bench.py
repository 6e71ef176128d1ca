"""Timing of the search kernels.  Usage: python tools/search_bench.py [--n 2000000] [--q 4096] [--k 10] [--iters 5]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_lora_match_b200 import _lib, kernels as K

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2_000_000); ap.add_argument("--q", type=int, nargs="+", default=[4096])
    ap.add_argument("--d", type=int, default=768); ap.add_argument("--k", type=int, default=10); ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(4)
    e = torch.randn((a.n, a.d), generator=g, device=dev); e = e / e.norm(dim=-1, keepdim=True)
    eb = e.bfloat16()
    lib = _lib.load()
    for nq in a.q:
        q = torch.randn((nq, a.d), generator=g, device=dev); q = q / q.norm(dim=-1, keepdim=True)
        qb = q.bfloat16()
        for _ in range(2): K.search_topk(q, qb, eb, e, a.k)
        lib.clm_prof_enable(1)
        for _ in range(a.iters): K.search_topk(q, qb, eb, e, a.k)
        ps, pm = _lib.prof_summary("search"), _lib.prof_summary("merge")
        recs = [round(r[3], 3) for r in _lib.prof_records() if r[0] in ("search", "merge")]
        print("  per-launch ms (sample scan, sample merge, main scan, final merge) x iters:", recs[:8], flush=True)
        lib.clm_prof_enable(0)
        ms = ps["ms"] / a.iters
        print(json.dumps({"n": a.n, "q": nq, "d": a.d, "k": a.k, "scan_ms": round(ms, 4), "merge_ms": round(pm["ms"] / a.iters, 4),
                          "tflops": round(ps["flops"] / a.iters / ms / 1e9, 1), "gbs": round(ps["bytes"] / a.iters / ms / 1e6, 1),
                          "qps": round(nq / ((ms + pm["ms"] / a.iters) / 1e3))}), flush=True)

if __name__ == "__main__":
    main()
