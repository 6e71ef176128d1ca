"""Per-launch timing of one LoRA training step (clm_prof_* events), grouped by (kind, algorithmic FLOPs, bytes):
which launches of the step cost what.  python tools/train_profile.py [--arch ...] [--batch 256]"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lora_match_b200 import _lib  # noqa: E402
from clip_lora_match_b200.models import clip_model as CM  # noqa: E402
from clip_lora_match_b200.models.lora_adapter import LoraConfig, init_lora_adapter  # noqa: E402
from clip_lora_match_b200.models.lora_trainer import LoraTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="openai/clip-vit-base-patch32")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--targets", default="q_proj,k_proj,v_proj,out_proj")
    ap.add_argument("--rank", type=int, default=8)
    ap.add_argument("--top", type=int, default=25)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    arch = CM.arch_from_name(a.arch)
    model = CM.B200ClipModel(arch, CM.random_init_state_dict(arch, seed=0), device=dev)
    model.set_lora(init_lora_adapter(model.linear_dims(), LoraConfig(r=a.rank, lora_alpha=2 * a.rank,
                                                                     target_modules=a.targets.split(",")), seed=1))
    tr = LoraTrainer(model, use_graph=False)
    pv = torch.randn((a.batch, 3, 224, 224), device=dev)
    ids = torch.randint(0, 49000, (a.batch, 77), device=dev, dtype=torch.int32)
    ids[:, 0], ids[:, 40:] = 49406, 49407
    for _ in range(3):
        tr.step(pv, ids)
    torch.cuda.synchronize()
    lib.clm_prof_enable(1)
    tr.step(pv, ids)
    recs = _lib.prof_records()
    lib.clm_prof_enable(0)
    groups = collections.defaultdict(lambda: [0, 0.0])
    for kind, flops, nbytes, ms in recs:
        g = groups[(kind, round(flops / 1e6), round(nbytes / 1e3))]
        g[0] += 1
        g[1] += ms
    total = sum(r[3] for r in recs)
    print(f"{a.arch} batch {a.batch}: {len(recs)} launches, kernel sum {total:.3f} ms")
    print(f"{'kind':12s} {'MFLOP':>10s} {'KB':>10s} {'n':>4s} {'ms':>8s} {'us each':>8s} {'TFLOP/s':>8s} {'GB/s':>8s}")
    for (kind, mf, kb), (n, ms) in sorted(groups.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print(f"{kind:12s} {mf:10d} {kb:10d} {n:4d} {ms:8.3f} {1e3 * ms / n:8.1f} {mf * n / ms / 1e3 if ms else 0:8.1f} "
              f"{kb * n / ms / 1e3 if ms else 0:8.0f}")


if __name__ == "__main__":
    main()
