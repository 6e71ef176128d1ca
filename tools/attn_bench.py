"""Timing of clm_attention at the BASELINE shapes.  Usage: python tools/attn_bench.py [--only b16]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_lora_match_b200 import kernels as K

SHAPES = {"b16": (1024, 197, 12, False), "l14": (1024, 257, 16, False), "b32": (1024, 50, 12, False),
          "txt512": (1024, 77, 8, True), "txt768": (1024, 77, 12, True)}

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--only", default=""); ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda")
    for name, (B, T, H, causal) in SHAPES.items():
        if a.only and a.only not in name: continue
        qkv = torch.randn((B * T, 3 * H * 64), device=dev).bfloat16()
        for _ in range(3): K.attention(qkv, B, T, H, causal)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters): K.attention(qkv, B, T, H, causal)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        flops = 4.0 * B * H * T * T * 64
        byts = 2.0 * B * T * 4 * H * 64
        print(json.dumps({"shape": name, "B": B, "T": T, "H": H, "causal": causal, "ms": round(ms, 4),
                          "tflops": round(flops / ms / 1e9, 1), "gbs": round(byts / ms / 1e6, 1)}), flush=True)

if __name__ == "__main__":
    main()
