"""ViT-L/14 image-tower step at batch 512 with the tower's micro-batch size swept (clm_encode_image walks the batch in
micro-batches that fit the caller's workspace): does keeping a micro-batch's activations inside the 126 MB L2 between
consecutive kernels pay for the smaller GEMM grids?  python tools/l14_microbatch.py [--arch ...] [--batch 512]"""
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
    ap.add_argument("--arch", default="openai/clip-vit-large-patch14")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--micro", default="512,256,128,64,32")
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda")
    arch = CM.arch_from_name(a.arch)
    model = CM.B200ClipModel(arch, CM.random_init_state_dict(arch, 0), device=dev)
    model.set_lora(init_lora_adapter(model.linear_dims(), LoraConfig(r=16, lora_alpha=32), seed=1, init_b_std=0.02))
    pv = torch.randn((a.batch, 3, 224, 224), device=dev)
    lib = model._lib
    for r in range(a.rounds):
        for mb in [int(x) for x in a.micro.split(",")]:
            model.max_workspace_bytes = lib.clm_tower_workspace_bytes(model._towers["vision"], mb)
            model._workspace = None
            for _ in range(3):
                model.encode_images(pv)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.steps):
                model.encode_images(pv)
                if i == a.steps // 2:
                    clk, pw = smi()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            print(json.dumps({"round": r, "micro_batch": mb, "ms_per_step": round(ms, 3),
                              "images_per_s": round(a.batch / ms * 1e3, 1), "sm_mhz": clk, "power_w": pw}), flush=True)


if __name__ == "__main__":
    main()
