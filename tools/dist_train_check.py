"""Data-parallel LoRA training on N GPUs against one process accumulating the same micro-batches.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_train_check.py
Every rank: LoraTrainer(distributed=True).step on its own micro-batch (one all-reduce of the flat gradient over
NCCL), then -- same process, fresh trainer -- grad_accum_steps = N over all N micro-batches.  The parameters after
the optimizer step must agree, and the DP step is timed."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lora_match_b200.models import clip_model as CM  # noqa: E402
from clip_lora_match_b200.models.lora_adapter import LoraConfig, init_lora_adapter  # noqa: E402
from clip_lora_match_b200.models.lora_trainer import LoraTrainer  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    batch = int(os.environ.get("CLM_TRAIN_BATCH", 8))
    arch = CM.arch_from_name("openai/clip-vit-base-patch32")
    sd = CM.random_init_state_dict(arch, seed=0)
    cfg = LoraConfig(r=8, lora_alpha=16, target_modules=["q_proj", "k_proj", "v_proj", "out_proj"])

    def make(**kw):
        m = CM.B200ClipModel(arch, sd, device=dev)
        m.set_lora(init_lora_adapter(m.linear_dims(), cfg, seed=1, init_b_std=0.02))
        return LoraTrainer(m, lr=1e-3, **kw)

    def data(r):
        g = torch.Generator(device=dev).manual_seed(100 + r)
        pv = torch.randn((batch, 3, 224, 224), generator=g, device=dev)
        ids = torch.randint(0, 49000, (batch, 77), generator=g, device=dev, dtype=torch.int32)
        ids[:, 0], ids[:, 20 + r:] = 49406, 49407
        return pv, ids

    dp = make(distributed=world > 1, deterministic=True)
    pv, ids = data(rank)
    loss_dp = dp.step(pv, ids).item()
    acc = make(grad_accum_steps=world, deterministic=True)
    total = 0.0
    for r in range(world):
        total += acc.forward_backward(*data(r)).item()
    acc.optimizer_step()
    upd_dp, upd_acc = dp.theta - make().theta, acc.theta - make().theta
    cos = float((upd_dp.double() @ upd_acc.double()) / (upd_dp.double().norm() * upd_acc.double().norm()))
    ok = abs(loss_dp - total) < 1e-3 and cos > 0.999
    for _ in range(5):
        dp.step(pv, ids)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    steps = 20
    for _ in range(steps):
        dp.step(pv, ids)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / steps
    print(f"[rank {rank}/{world}] DP loss {loss_dp:.5f} vs accumulated {total:.5f}; update cosine {cos:.6f}; "
          f"{ms:.3f} ms per DP step of {batch} pairs per GPU = {batch * world / ms * 1e3:.0f} pairs/s; {'OK' if ok else 'MISMATCH'}",
          flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
