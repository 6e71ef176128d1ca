"""Per-shape timing of clm_gemm_epi (and torch.matmul/cuBLAS as a same-box reference point).
Usage: python tools/gemm_bench.py [--ref] [--shapes vision|all] [--iters N]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_lora_match_b200 import kernels as K

SHAPES = {
    "b16_qkv": (201728, 2304, 768), "b16_out": (201728, 768, 768), "b16_fc1": (201728, 3072, 768),
    "b16_fc2": (201728, 768, 3072), "b16_lora_down": (201728, 64, 768),
    "l14_qkv": (263168, 3072, 1024), "l14_fc1": (263168, 4096, 1024), "l14_fc2": (263168, 1024, 4096),
    "txt_qkv": (78848, 1536, 512), "txt_fc1": (78848, 2048, 512), "txt_fc2": (78848, 512, 2048),
    "square_8k": (8192, 8192, 8192),
}

def bench(fn, iters):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", action="store_true"); ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda")
    for name, (M, N, Kd) in SHAPES.items():
        if args.only and args.only not in name: continue
        a = torch.randn((M, Kd), device=dev).bfloat16(); w = torch.randn((N, Kd), device=dev).bfloat16()
        bias = torch.randn((N,), device=dev)
        out = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
        ms = bench(lambda: K.gemm_epi(a, w, bias=bias, out=out), args.iters)
        row = {"shape": name, "M": M, "N": N, "K": Kd, "clm_ms": round(ms, 4), "clm_tflops": round(2 * M * N * Kd / ms / 1e9, 1)}
        if args.ref:
            wt = w.t()
            ms2 = bench(lambda: torch.matmul(a, wt, out=out), args.iters)
            row["cublas_ms"] = round(ms2, 4); row["cublas_tflops"] = round(2 * M * N * Kd / ms2 / 1e9, 1)
        print(json.dumps(row), flush=True)

if __name__ == "__main__":
    main()
