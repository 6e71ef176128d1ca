"""Launch list of one encoder step (events around every launch), grouped by (kind, flops) signature.
Usage: python tools/step_profile.py [--arch openai/clip-vit-base-patch16] [--batch 1024] [--tower image|text|both]"""
import argparse, collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_lora_match_b200 import _lib
from clip_lora_match_b200.models import clip_model as CM
from clip_lora_match_b200.models.lora_adapter import LoraConfig, init_lora_adapter

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="openai/clip-vit-base-patch16"); ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--tower", default="both"); ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda")
    arch = CM.arch_from_name(a.arch)
    model = CM.B200ClipModel(arch, CM.random_init_state_dict(arch, 0), device=dev)
    model.set_lora(init_lora_adapter(model.linear_dims(), LoraConfig(r=16, lora_alpha=32), seed=1, init_b_std=0.02))
    pv = torch.randn((a.batch, 3, 224, 224), device=dev)
    ids = torch.randint(0, 49405, (a.batch, 77), device=dev, dtype=torch.int32); ids[:, 0] = 49406; ids[:, 40:] = 49407
    def step():
        if a.tower in ("both", "image"): model.encode_images(pv)
        if a.tower in ("both", "text"): model.encode_texts(ids)
    for _ in range(3): step()
    lib = _lib.load()
    lib.clm_prof_enable(1)
    for _ in range(a.reps): step()
    recs = _lib.prof_records()
    lib.clm_prof_enable(0)
    groups = collections.OrderedDict()
    for kind, fl, by, ms in recs:
        g = groups.setdefault((kind, round(fl), round(by)), [0, 0.0])
        g[0] += 1; g[1] += ms
    total = sum(g[1] for g in groups.values())
    print(f"# {a.arch} batch {a.batch} tower {a.tower}: {len(recs)//a.reps} launches/step, {total/a.reps:.3f} ms/step kernel time")
    for (kind, fl, by), (n, ms) in groups.items():
        avg = ms / n
        print(json.dumps({"kind": kind, "launches_per_step": n // a.reps, "avg_ms": round(avg, 4), "ms_per_step": round(ms / a.reps, 3),
                          "share": round(ms / total, 4), "gflop": round(fl / 1e9, 2), "tflops": round(fl / avg / 1e9, 1) if fl else None,
                          "mbytes": round(by / 1e6, 1), "gbs": round(by / avg / 1e6, 1)}))

if __name__ == "__main__":
    main()
