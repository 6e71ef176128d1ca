"""Same-process A/B of the residual stream's type (fp32 vs bf16, B200ClipModel.set_residual_dtype): image-tower step
of ViT-L/14 at batch 512 and of ViT-B/16 at batch 1024, unprofiled (CUDA events around --steps steps), then one
profiled pass per arm for the per-kind kernel sums, and the cosine between the two arms' embeddings.
Usage: python tools/residual_ab.py [--rounds 3] [--steps 10] [--archs l14,b16]"""
import argparse, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_lora_match_b200 import _lib
from clip_lora_match_b200.models import clip_model as CM
from clip_lora_match_b200.models.lora_adapter import LoraConfig, init_lora_adapter

ARMS = (("float32", False), ("bfloat16", False), ("bfloat16", True))
ARCHS = {"l14": ("openai/clip-vit-large-patch14", 512), "b16": ("openai/clip-vit-base-patch16", 1024)}


def smi():
    try:
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        return float(o[0]), float(o[1])
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=3); ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--archs", default="l14,b16")
    a = ap.parse_args()
    dev = torch.device("cuda")
    lib = _lib.load()
    for key in a.archs.split(","):
        name, batch = ARCHS[key]
        arch = CM.arch_from_name(name)
        model = CM.B200ClipModel(arch, CM.random_init_state_dict(arch, 0), device=dev)
        model.set_lora(init_lora_adapter(model.linear_dims(), LoraConfig(r=16, lora_alpha=32), seed=1, init_b_std=0.02))
        pv = torch.randn((batch, 3, 224, 224), device=dev)
        emb = {}
        for _ in range(3): model.encode_images(pv)
        for r in range(a.rounds):
            for dt, fold in ARMS:
                model.set_residual_dtype(dt); model.set_ln_fold(fold)
                dt = dt + ("+fold" if fold and dt == "bfloat16" else "")
                for _ in range(3): emb[dt] = model.encode_images(pv)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                clk = pw = None
                for i in range(a.steps):
                    model.encode_images(pv)
                    if i == a.steps // 2: clk, pw = smi()
                e1.record(); torch.cuda.synchronize()
                print(json.dumps({"arch": key, "round": r, "residual": dt, "ms_per_step": round(e0.elapsed_time(e1) / a.steps, 3),
                                  "sm_mhz": clk, "power_w": pw}), flush=True)
        for dt, fold in ARMS:
            model.set_residual_dtype(dt); model.set_ln_fold(fold)
            dt = dt + ("+fold" if fold and dt == "bfloat16" else "")
            for _ in range(2): model.encode_images(pv)
            lib.clm_prof_enable(1)
            for _ in range(2): model.encode_images(pv)
            recs = _lib.prof_records()
            lib.clm_prof_enable(0)
            recs = recs[-(len(recs) // 2):]
            kinds = {k: round(sum(x[3] for x in recs if x[0] == k), 3) for k in ("gemm", "attention", "elementwise")}
            print(json.dumps({"arch": key, "residual": dt, "kernel_ms_profiled_step": kinds, "launches": len(recs)}), flush=True)
        for other in ("bfloat16", "bfloat16+fold"):
            cos = torch.nn.functional.cosine_similarity(emb["float32"], emb[other], dim=-1)
            print(json.dumps({"arch": key, "cos_vs_fp32_stream": other, "min": float(cos.min()), "mean": float(cos.mean())}), flush=True)
        del model, pv
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
