"""Every launch of one search_topk call (sample GEMM, k-th largest, scan, merge) and the call's wall time.
Usage: python tools/search_profile.py [--n 1250000] [--q 4096] [--k 10 50]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_lora_match_b200 import _lib, kernels as K

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_250_000); ap.add_argument("--q", type=int, default=4096)
    ap.add_argument("--d", type=int, default=768); ap.add_argument("--k", type=int, nargs="+", default=[10, 50])
    ap.add_argument("--list-cap", type=int, default=0); ap.add_argument("--margin", type=float, default=-1.0); ap.add_argument("--sample-rows", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(4)
    e = torch.randn((a.n, a.d), generator=g, device=dev); e = e / e.norm(dim=-1, keepdim=True)
    eb = e.bfloat16()
    q = torch.randn((a.q, a.d), generator=g, device=dev); q = q / q.norm(dim=-1, keepdim=True)
    qb = q.bfloat16()
    lib = _lib.load()
    for k in a.k:
        kw = {"list_cap": a.list_cap} if a.list_cap else {}
        if a.margin >= 0: kw["margin"] = a.margin
        if a.sample_rows: kw["sample_rows"] = a.sample_rows
        st = {}
        for _ in range(3): K.search_topk(q, qb, eb, e, k, stats=st, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): K.search_topk(q, qb, eb, e, k, **kw)
        e1.record(); torch.cuda.synchronize()
        wall = e0.elapsed_time(e1) / 5
        lib.clm_prof_enable(1)
        K.search_topk(q, qb, eb, e, k, **kw)
        recs = [(r[0], round(r[3], 4)) for r in _lib.prof_records()]
        lib.clm_prof_enable(0)
        print(json.dumps({"n": a.n, "q": a.q, "k": k, "wall_ms_per_call": round(wall, 3), "stats": st, "launches": recs}), flush=True)

if __name__ == "__main__":
    main()
