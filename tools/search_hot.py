"""Trajectory of the streaming scan (Q <= 128) from a cold board into the power-capped state.
Usage: python tools/search_hot.py [--n 10000000] [--q 64] [--blocks 20] [--per 20] [--preheat 0]
Prints the mean scan time of consecutive blocks of `per` searches (kernel time from the library's
launch profiler) together with the SM clock / power sampled by nvidia-smi during the block."""
import argparse, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_lora_match_b200 import _lib, kernels as K

def smi():
    try:
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        return float(o[0]), float(o[1])
    except Exception:
        return None, None

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000); ap.add_argument("--q", type=int, default=64)
    ap.add_argument("--d", type=int, default=768); ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--blocks", type=int, default=20); ap.add_argument("--per", type=int, default=20)
    ap.add_argument("--preheat", type=int, default=0, help="seconds of bf16 matmul before the scans (tensor-bound phase)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(4)
    e = torch.empty((a.n, a.d), device=dev)
    for s0 in range(0, a.n, 1_000_000):
        blk = torch.randn((min(1_000_000, a.n - s0), a.d), generator=g, device=dev)
        e[s0:s0 + blk.shape[0]] = blk / blk.norm(dim=-1, keepdim=True)
    eb = e.bfloat16()
    q = torch.randn((a.q, a.d), generator=g, device=dev); q = q / q.norm(dim=-1, keepdim=True)
    qb = q.bfloat16()
    lib = _lib.load()
    for _ in range(2): K.search_topk(q, qb, eb, e, a.k)
    torch.cuda.synchronize()
    if a.preheat:
        x = torch.randn((8192, 8192), device=dev).bfloat16()
        import time
        t0 = time.time()
        while time.time() - t0 < a.preheat:
            for _ in range(20): x @ x
            torch.cuda.synchronize()
    out = []
    for b in range(a.blocks):
        lib.clm_prof_enable(1)
        for _ in range(a.per): K.search_topk(q, qb, eb, e, a.k)
        mhz, w = smi()
        main_ms = sorted(r[3] for r in _lib.prof_records() if r[0] == "search")[-a.per:]  # the main scans (sample scans are tiny)
        lib.clm_prof_enable(0)
        ms = sum(main_ms) / len(main_ms)
        out.append({"block": b, "scan_ms": round(ms, 4), "gbs": round(2.0 * a.d * (a.n + a.q) / ms / 1e6, 1), "sm_mhz": mhz, "power_w": w})
        print(json.dumps(out[-1]), flush=True)

if __name__ == "__main__":
    main()
