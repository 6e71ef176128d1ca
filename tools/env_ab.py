"""Same-process A/B of an environment toggle the library reads on every call: image-tower step of ViT-L/14 (batch 512)
and ViT-B/16 (batch 1024), CUDA events around --steps steps, arms interleaved.
Usage: python tools/env_ab.py --env CLM_GEMM_DIRECT_STORE [--a 0 --b 1] [--rounds 3] [--archs l14,b16]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_lora_match_b200 import _lib
from clip_lora_match_b200.models import clip_model as CM
from clip_lora_match_b200.models.lora_adapter import LoraConfig, init_lora_adapter

ARCHS = {"l14": ("openai/clip-vit-large-patch14", 512), "b16": ("openai/clip-vit-base-patch16", 1024)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", required=True); ap.add_argument("--a", default="0"); ap.add_argument("--b", default="1")
    ap.add_argument("--rounds", type=int, default=3); ap.add_argument("--steps", type=int, default=8)
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
            for val in (a.a, a.b):
                os.environ[a.env] = val
                for _ in range(2): emb[val] = model.encode_images(pv)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.steps): model.encode_images(pv)
                e1.record(); torch.cuda.synchronize()
                print(json.dumps({"arch": key, "round": r, a.env: val, "ms_per_step": round(e0.elapsed_time(e1) / a.steps, 3)}), flush=True)
        for val in (a.a, a.b):
            os.environ[a.env] = val
            for _ in range(2): model.encode_images(pv)
            lib.clm_prof_enable(1)
            for _ in range(2): model.encode_images(pv)
            recs = _lib.prof_records()
            lib.clm_prof_enable(0)
            recs = recs[-(len(recs) // 2):]
            kinds = {k: round(sum(x[3] for x in recs if x[0] == k), 3) for k in ("gemm", "attention", "elementwise")}
            print(json.dumps({"arch": key, a.env: val, "kernel_ms_profiled_step": kinds}), flush=True)
        print(json.dumps({"arch": key, "max_abs_diff_between_arms": float((emb[a.a] - emb[a.b]).abs().max())}), flush=True)
        del model, pv
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
