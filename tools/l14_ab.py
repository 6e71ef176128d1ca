"""Same-process A/B of the ViT-L/14 image-tower step (batch 512, unprofiled, CUDA events around 10 steps) with an
environment toggle that the library reads on every call (CLM_ATTN_SPLIT).  Prints ms/step, SM clock and power per arm.
Usage: python tools/l14_ab.py [--env CLM_ATTN_SPLIT] [--a 0] [--b 1] [--rounds 3]"""
import argparse, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_lora_match_b200.models import clip_model as CM
from clip_lora_match_b200.models.lora_adapter import LoraConfig, init_lora_adapter

def smi():
    try:
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        return float(o[0]), float(o[1])
    except Exception:
        return None, None

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="CLM_ATTN_SPLIT"); ap.add_argument("--a", default="0"); ap.add_argument("--b", default="1")
    ap.add_argument("--rounds", type=int, default=3); ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--arch", default="openai/clip-vit-large-patch14"); ap.add_argument("--batch", type=int, default=512)
    a = ap.parse_args()
    dev = torch.device("cuda")
    arch = CM.arch_from_name(a.arch)
    model = CM.B200ClipModel(arch, CM.random_init_state_dict(arch, 0), device=dev)
    model.set_lora(init_lora_adapter(model.linear_dims(), LoraConfig(r=16, lora_alpha=32), seed=1, init_b_std=0.02))
    pv = torch.randn((a.batch, 3, 224, 224), device=dev)
    for _ in range(5): model.encode_images(pv)
    for r in range(a.rounds):
        for val in (a.a, a.b):
            os.environ[a.env] = val
            for _ in range(2): model.encode_images(pv)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.steps):
                model.encode_images(pv)
                if i == a.steps // 2: clk, pw = smi()
            e1.record(); torch.cuda.synchronize()
            print(json.dumps({"round": r, a.env: val, "ms_per_step": round(e0.elapsed_time(e1) / a.steps, 3), "sm_mhz": clk, "power_w": pw}), flush=True)

if __name__ == "__main__":
    main()
