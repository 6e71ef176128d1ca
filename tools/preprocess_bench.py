"""Throughput of clm_preprocess_images.  Usage: python tools/preprocess_bench.py [--batch 256] [--h 480] [--w 640]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_lora_match_b200 import _lib, kernels as K

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--sizes", type=int, nargs="+", default=[480, 640, 1080, 1920, 300, 260])
    a = ap.parse_args()
    dev = torch.device("cuda")
    for i in range(0, len(a.sizes), 2):
        h, w = a.sizes[i], a.sizes[i + 1]
        imgs = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device=dev) for _ in range(a.batch)]
        out = torch.empty((a.batch, 3, 224, 224), device=dev)
        for _ in range(3): K.preprocess_images(imgs, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): K.preprocess_images(imgs, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        byts = a.batch * (3.0 * h * w + 12.0 * 224 * 224)
        print(json.dumps({"h": h, "w": w, "batch": a.batch, "ms": round(ms, 3), "images_per_s": round(a.batch / ms * 1e3),
                          "gbs": round(byts / ms / 1e6, 1)}), flush=True)

if __name__ == "__main__":
    main()
