#!/usr/bin/env python
"""bench.py — the retrieval hot path on N B200s of one node (contract: see DESIGN.md §Measurement).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle) on the host cores

Workload (BASELINE.json configs[1]): CLIP ViT-B/16 + LoRA r=16 (q,v), one step = 1024 synthetic
224px images + 1024 synthetic 77-token captions through both towers, random-init weights.
`value` is image-caption pairs per second (each image is encoded together with its caption; the
image-tower-only and text-tower-only rates are in `detail`).  The search half of the metric
(top-10 over a 10M x 768 index, 4096-query batches, row-sharded over the N GPUs) is reported
under `search` in the same JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ARCH = "openai/clip-vit-base-patch16"
LORA_R, LORA_ALPHA, LORA_TARGETS = 16, 32, ["q_proj", "v_proj"]
BATCH = 1024
L14_ARCH, L14_BATCH = "openai/clip-vit-large-patch14", 512
TRAIN_ARCH, TRAIN_TARGETS, TRAIN_BATCHES = "openai/clip-vit-base-patch32", ["q_proj", "k_proj", "v_proj", "out_proj"], (8, 256)
INDEX_ROWS, INDEX_DIM, QUERY_BATCH, TOP_K = 10_000_000, 768, 4096, 10


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"],
                "tf_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


def synth_captions(batch, seed=3, context=77, bos=49406, eos=49407):
    """SURVEY.md §8(d) captions: ids [B,77] = [BOS, tok..., EOS, EOS...], tok ~ U[0,49405], length ~ U{3..77};
    mask = pos < length.  (Same generator as the oracle's, restated here so the B200 arm imports nothing
    from oracle/.)"""
    import torch

    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(3, context + 1, (batch,), generator=g)
    ids = torch.randint(0, bos, (batch, context), generator=g)
    pos = torch.arange(context).unsqueeze(0)
    ids[:, 0] = bos
    ids = torch.where(pos >= (lengths - 1).unsqueeze(1), torch.full_like(ids, eos), ids)
    return ids, (pos < lengths.unsqueeze(1)).long()


def flops_per_item(T, D, L, M, P, r, n_lora, patch_k=0, np_=0):
    """L(8TD^2 + 4TDM + 4T^2D + n_lora 4TDr) + 2 Np 3P^2 D + 2DP."""
    return L * (8 * T * D * D + 4 * T * D * M + 4 * T * T * D + n_lora * 4 * T * D * r) + \
        2 * np_ * patch_k * D + 2 * D * P


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        clocks, maxc, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                clocks.append(float(r[1])); maxc.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not clocks:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples drawing more than half of the highest power seen
        hot = [c for c, p in zip(clocks, power) if p >= 0.5 * max(power)] or clocks
        return {"sm_mhz": statistics.median(hot), "sm_max_mhz": max(maxc), "power_w_max": max(power),
                "samples": len(clocks), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own path (oracle port) on the host cores
# ------------------------------------------------------------------------------------------
def cpu_encode_rate(target_seconds=12.0, per_step=8):
    """Oracle (transformers fp32 CLIPModel + PEFT-semantics LoRA + reference post-processing) on
    all host threads: image+caption pairs per second on a bounded sample of the same workload."""
    import torch

    from oracle import clip_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = O.build_model(ARCH, seed=0)
    O.synthetic_lora(model, LORA_R, LORA_ALPHA, LORA_TARGETS, seed=1)
    pv = O.synth_images(per_step, seed=2)
    ids, mask = O.synth_captions(per_step, seed=3)
    O.encode_images(model, pv[:2]); O.encode_texts(model, ids[:2], mask[:2])  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        O.encode_images(model, pv, batch_size=16)
        O.encode_texts(model, ids, mask, batch_size=16)
        n += per_step
        dt = time.perf_counter() - t0
        if dt >= target_seconds:
            break
    return n / dt, cores, f"{n} image+caption pairs of configs[1] (ViT-B/16+LoRA r=16, fp32, batch {per_step}) in {dt:.1f}s"


def cpu_train_rate(target_seconds=10.0, batch=8):
    """The training oracle (transformers fp32 CLIPModel + LoRA wrappers, torch autograd, torch AdamW: the reference's
    step loop, scripts/train_lora.py:172-193) on all host threads: pairs per second over ~10 s of optimizer steps at
    the reference's own batch of 8."""
    import torch

    from oracle import clip_oracle as O
    from oracle import train_oracle as T

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = O.build_model(TRAIN_ARCH, seed=0)
    O.synthetic_lora(model, 8, 16, TRAIN_TARGETS, seed=1, b_std=0.0)
    oracle = T.TrainOracle(model, lr=1e-4)
    pv = O.synth_images(batch, seed=8)
    ids, mask = O.synth_captions(batch, seed=9)
    oracle.step(pv, ids, mask)  # warm-up
    steps, t0 = 0, time.perf_counter()
    while True:
        oracle.step(pv, ids, mask)
        steps += 1
        dt = time.perf_counter() - t0
        if dt >= target_seconds:
            break
    return batch * steps / dt, cores, f"{steps} optimizer steps of batch {batch} (ViT-B/32 + LoRA r=8 q,k,v,out, fp32 autograd) in {dt:.1f}s"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import clip_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = O.build_model(ARCH, seed=0)
    O.synthetic_lora(model, LORA_R, LORA_ALPHA, LORA_TARGETS, seed=1)
    per_step = 8
    pv = O.synth_images(per_step, seed=2)
    ids, mask = O.synth_captions(per_step, seed=3)

    def step():
        O.encode_images(model, pv, batch_size=16)
        O.encode_texts(model, ids, mask, batch_size=16)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"each step = {per_step} image+caption pairs of configs[1] (bounded sample of the 1024-pair step)"
    line = {
        "impl": "reference", "metric": "CLIP+LoRA images/sec", "value": value, "unit": "image-caption pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "image-caption pairs/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "image-caption pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "configs[1]: CLIP ViT-B/16 + LoRA r=16 (q,v) image+text encoding, 1024 images + "
                        "1024 captions per step per GPU, random-init weights",
            "arch": ARCH, "lora": {"r": LORA_R, "alpha": LORA_ALPHA, "targets": LORA_TARGETS, "merged": False},
            "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus,
            "tokens": {"image": 197, "text": 77},
            "parallelism": f"dp{n_gpus}", "l2": "inputs larger than L2 (616 MB pixel_values per step)"}


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-search", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l14", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--index-rows", type=int, default=INDEX_ROWS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from clip_lora_match_b200 import _lib
    from clip_lora_match_b200.models import clip_model as CM
    from clip_lora_match_b200.models.lora_adapter import LoraConfig, init_lora_adapter
    from clip_lora_match_b200.src.embedding.search import TextSearchIndex, shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime

        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    lib = _lib.load()
    peaks = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps):
        """K steps bracketed by barrier+synchronize, CUDA events on the launching stream, max over ranks."""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b))

    # ---- model + inputs (resident in HBM before the timed region) ------------------------
    arch = CM.arch_from_name(ARCH)
    sd = CM.random_init_state_dict(arch, seed=0)
    model = CM.B200ClipModel(arch, sd, device=dev)
    lora = init_lora_adapter(model.linear_dims(), LoraConfig(r=LORA_R, lora_alpha=LORA_ALPHA,
                                                             target_modules=LORA_TARGETS),
                             seed=1, init_b_std=0.02, base_model_name=ARCH)
    model.set_lora(lora)
    g = torch.Generator(device=dev).manual_seed(2 + rank)
    pv = torch.randn((BATCH, 3, 224, 224), generator=g, device=dev)
    ids_cpu, mask_cpu = synth_captions(BATCH, seed=3 + rank)
    ids = ids_cpu.to(dev, torch.int32)
    # Every caption is encoded on all 77 positions (SURVEY.md §8d counts a caption as a 77-token problem), so
    # no position is skipped in the timed region.  CLM_BENCH_BUCKETED_TEXT=1 runs the length-bucketed passes
    # of encode_texts instead (caption lengths U{3..77}); measured, that is a wash on this length mix.
    bucketed = os.environ.get("CLM_BENCH_BUCKETED_TEXT") == "1"
    cap_len = mask_cpu.sum(dim=1) if bucketed else None

    out = {}

    def step():
        out["img"] = model.encode_images(pv)
        out["txt"] = model.encode_texts(ids, lengths=cap_len)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.clm_launch_count()
    ms = timed(step, args.steps)
    launches = lib.clm_launch_count() - l0
    ms_img = timed(lambda: model.encode_images(pv), args.steps)
    ms_txt = timed(lambda: model.encode_texts(ids, lengths=cap_len), args.steps)
    clocks = sampler.stop()
    assert torch.isfinite(out["img"]).all() and torch.isfinite(out["txt"]).all()
    value = BATCH * world * args.steps / (ms / 1e3)

    # ---- per-kernel durations, measured live with events around every launch -----------
    # The profiled pass brackets every launch with two CUDA events on the launching stream.  Its accounting
    # (sum of kernel durations, sum of the gaps between consecutive launches, its own wall time) is reported
    # beside the unprofiled step time so that the difference between ms_per_step and the kernel sum is explained
    # rather than assumed: see detail.step_accounting.
    # Three profiled steps follow two unprofiled ones without any host synchronisation in between, and only the
    # LAST step's records are used: a single step launched after a synchronize starts with power headroom and
    # runs ~10 % faster than the sustained state the timed region is in (measured: 46.3 vs 50.9 ms).
    PROF_STEPS = 3
    for _ in range(2):
        step()
    lib.clm_prof_enable(1)
    pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pa.record()
    for _ in range(PROF_STEPS):
        step()
    pb.record()
    torch.cuda.synchronize()
    prof_step_ms = pa.elapsed_time(pb) / PROF_STEPS
    recs_all = _lib.prof_records()
    tl_all = _lib.prof_timeline()
    lib.clm_prof_enable(0)
    n_step = len(recs_all) // PROF_STEPS
    recs, tl = recs_all[-n_step:], tl_all[-n_step:]
    prof = {}
    for kname in ("gemm", "attention", "elementwise"):
        sel = [r for r in recs if r[0] == kname]
        prof[kname] = {"ms": sum(r[3] for r in sel), "flops": sum(r[1] for r in sel), "bytes": sum(r[2] for r in sel),
                       "launches": len(sel)}
    gemm = prof["gemm"]
    step_kernel_ms = sum(p["ms"] for p in prof.values())
    gaps = [tl[i + 1][1] - (tl[i][1] + tl[i][2]) for i in range(len(tl) - 1)]
    big_gaps = sorted(((g, i) for i, g in enumerate(gaps)), reverse=True)[:3]
    accounting = {
        "ms_per_step_unprofiled": ms / args.steps,
        "profiled_step_wall_ms": prof_step_ms,
        "sum_kernel_ms_profiled": step_kernel_ms,
        "sum_gaps_between_launches_ms_profiled": sum(g for g in gaps if g > 0),
        "launches": len(tl),
        "largest_gaps_ms": [{"after_launch": i, "kind_before": tl[i][0], "kind_after": tl[i + 1][0], "gap_ms": round(g, 4)}
                            for g, i in big_gaps],
        "note": "kernel durations are those of the last of three back-to-back profiled steps (two events around every "
                "launch) that follow unprofiled steps without a host synchronisation, i.e. of the sustained power-capped "
                "state; a single profiled step launched after a synchronize ran ~10 % faster (boost clocks), which is what "
                "made round 1's kernel sum fall short of ms_per_step",
    }
    # dominant kernel TYPE: every gemm_kernel launch of the step (77 % of its kernel time); the fc1 launches
    # (largest single shape) are reported as a sub-entry
    fc1_flops = 2.0 * BATCH * 197 * arch.vision.mlp * arch.vision.width
    fc1_bytes = 2.0 * (BATCH * 197 * arch.vision.width + arch.vision.mlp * arch.vision.width) + \
        2.0 * BATCH * 197 * arch.vision.mlp  # fc2 has the same FLOPs but other bytes
    fc1 = [r for r in recs if r[0] == "gemm" and abs(r[1] - fc1_flops) < 1e-3 * fc1_flops
           and abs(r[2] - fc1_bytes) < 1e-3 * fc1_bytes]
    fc1_ms = sum(r[3] for r in fc1) / max(1, len(fc1))
    fc1_tf = fc1_flops / (fc1_ms * 1e-3) / 1e12 if fc1 else 0.0
    all_tf = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12
    traffic = None
    for tname in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("traffic_bytes_per_launch")
            break
    n_gemm = max(1, gemm["launches"])
    roofline = {"bound": "tensor",
                "kernel": "gemm_kernel<*>: all tcgen05 GEMM launches of the step (QKV + LoRA, out-proj, fc1 + QuickGELU, "
                          "fc2, patch embed, projections; cta_group::2 256x256 tiles for the large ones)",
                "achieved": all_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": all_tf / peaks["tf_sustained"], "frac_of_burst": all_tf / peaks["tf_burst"],
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernels timed inside a long step)",
                "algorithmic_flops_per_launch": gemm["flops"] / n_gemm, "algorithmic_bytes_per_launch": gemm["bytes"] / n_gemm,
                "avg_launch_ms": gemm["ms"] / n_gemm, "launches_per_step": gemm["launches"],
                "share_of_step_kernel_time": gemm["ms"] / step_kernel_ms,
                "traffic": traffic,
                "traffic_source": "ncu --set full of the vision fc1 launch (bytes per launch; profiles/), the largest single shape",
                "fc1_launches": {"kernel": "gemm_kernel<256,2,store_bf16>: vision fc1 + QuickGELU, M=201728 N=3072 K=768",
                                 "achieved": fc1_tf, "frac": fc1_tf / peaks["tf_sustained"],
                                 "frac_of_burst": fc1_tf / peaks["tf_burst"], "avg_launch_ms": fc1_ms,
                                 "launches_per_step": len(fc1), "algorithmic_flops_per_launch": fc1_flops,
                                 "algorithmic_bytes_per_launch": fc1_bytes,
                                 "share_of_step_kernel_time": sum(r[3] for r in fc1) / step_kernel_ms}}
    va, ta = arch.vision, arch.text
    fl_img = flops_per_item(197, va.width, va.layers, va.mlp, arch.proj_dim, LORA_R, 2, 3 * 16 * 16, 196)
    fl_txt = flops_per_item(77, ta.width, ta.layers, ta.mlp, arch.proj_dim, LORA_R, 2)
    if bucketed:  # positions after a caption's EOS are skipped: count what a caption of its true length costs
        fl_txt = sum(flops_per_item(int(L), ta.width, ta.layers, ta.mlp, arch.proj_dim, LORA_R, 2)
                     for L in mask_cpu.sum(dim=1).tolist()) / BATCH
    step_tf = (fl_img + fl_txt) * BATCH / (ms / args.steps * 1e-3) / 1e12
    detail = {
        "images_per_s_image_tower_only": BATCH * world * args.steps / (ms_img / 1e3),
        "texts_per_s_text_tower_only": BATCH * world * args.steps / (ms_txt / 1e3),
        "algorithmic_gflop_per_image": fl_img / 1e9, "algorithmic_gflop_per_caption": fl_txt / 1e9,
        "text_batching": "length-bucketed passes" if bucketed else "one 77-position pass (no position skipped)",
        "whole_step_tflops_per_gpu": step_tf, "whole_step_frac_of_sustained_peak": step_tf / peaks["tf_sustained"],
        "whole_step_frac_of_burst_peak": step_tf / peaks["tf_burst"],
        "kernel_ms_per_step": {k: {"ms": round(v["ms"], 3), "launches": v["launches"]} for k, v in prof.items()},
        "step_accounting": accounting,
    }

    # ---- the reference's own calling pattern: one image + one caption at a time (launch bound; the
    # host mirror replays a CUDA graph per batch size for small batches) ------------------------------
    pv1, ids1 = pv[:1].contiguous(), ids[:1].contiguous()

    def one_pair():
        model.encode_images(pv1)
        model.encode_texts(ids1)

    for _ in range(5):
        one_pair()
    ms_one = timed(one_pair, 50) / 50
    detail["single_pair_latency_ms"] = ms_one
    detail["single_pair_note"] = "batch-1 image + caption through encode_images/encode_texts (CUDA-graph replay), device inputs"

    # ---- the same step with the reference's fp32 residual stream (the headline runs the bf16 stream) -----------
    stream_default = model.residual_dtype
    detail["residual_stream"] = stream_default + (" (layer_norm1 / layer_norm2 folded into the QKV / fc1 GEMMs)"
                                                  if stream_default == "bfloat16" and model.ln_fold else "")
    if stream_default != "float32":
        model.set_residual_dtype("float32")
        for _ in range(args.warmup):
            step()
        ms32 = timed(step, args.steps)
        model.set_residual_dtype(stream_default)
        for _ in range(2):
            step()
        detail["fp32_residual_stream"] = {
            "ms_per_step": ms32 / args.steps, "pairs_per_s": BATCH * world * args.steps / (ms32 / 1e3),
            "note": "same step, same timing rules, residual stream between the layers in fp32 with standalone LayerNorm "
                    "passes (B200ClipModel(residual_dtype='float32')); the headline's bf16 stream is within cosine "
                    "0.9999 of it (tests/test_residual_bf16_gpu.py)"}

    # ---- end to end: host (pinned) inputs -> embeddings back on the host, every step ----
    pv_host = pv.cpu().pin_memory()
    ids_host = ids_cpu.to(torch.int32).pin_memory()
    emb_host = torch.empty((2, BATCH, arch.proj_dim), dtype=torch.float32).pin_memory()

    def e2e_step():
        img = model.encode_images(pv_host)  # pinned host tensor: chunked H2D overlapped with the encoder
        txt = model.encode_texts(ids_host, bucket=bucketed)
        emb_host[0].copy_(img, non_blocking=True)
        emb_host[1].copy_(txt, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e = {"value": BATCH * world * args.steps / (ms_e2e / 1e3), "unit": "image-caption pairs/s",
           "h2d_bytes_per_step": pv_host.numel() * 4 + ids_host.numel() * 4,
           "d2h_bytes_per_step": emb_host.numel() * 4, "ms_per_step": ms_e2e / args.steps}
    del pv_host, model, pv
    torch.cuda.empty_cache()

    # ---- ViT-L/14 + LoRA r=16 image tower (BASELINE configs[2] / north_star's >= 70 % target) -------
    # Reported beside the headline: images/s of the index-build encoder at batch 512 per GPU, and its
    # algorithmic FLOP rate against the measured bf16 peaks.  Same timing rules as the headline.
    l14 = None
    if not args.no_l14:
        arch_l = CM.arch_from_name(L14_ARCH)
        model_l = CM.B200ClipModel(arch_l, CM.random_init_state_dict(arch_l, seed=0), device=dev)
        model_l.set_lora(init_lora_adapter(model_l.linear_dims(), LoraConfig(r=LORA_R, lora_alpha=LORA_ALPHA,
                                                                             target_modules=LORA_TARGETS),
                                           seed=1, init_b_std=0.02, base_model_name=L14_ARCH))
        gl = torch.Generator(device=dev).manual_seed(6 + rank)
        pv_l = torch.randn((L14_BATCH, 3, 224, 224), generator=gl, device=dev)
        for _ in range(args.warmup):
            model_l.encode_images(pv_l)
        ms_l = timed(lambda: model_l.encode_images(pv_l), args.steps)
        lv = arch_l.vision
        fl_l = flops_per_item(257, lv.width, lv.layers, lv.mlp, arch_l.proj_dim, LORA_R, 2, 3 * 14 * 14, 256)
        tf_l = fl_l * L14_BATCH / (ms_l / args.steps * 1e-3) / 1e12
        for _ in range(2):  # sustained state: see the configs[1] profile above
            model_l.encode_images(pv_l)
        lib.clm_prof_enable(1)
        for _ in range(2):
            model_l.encode_images(pv_l)
        recs_l = _lib.prof_records()
        lib.clm_prof_enable(0)
        recs_l = recs_l[-(len(recs_l) // 2):]
        prof_l = {kn: {"ms": sum(r[3] for r in recs_l if r[0] == kn), "flops": sum(r[1] for r in recs_l if r[0] == kn),
                       "launches": sum(1 for r in recs_l if r[0] == kn)} for kn in ("gemm", "attention", "elementwise")}
        l14 = {"workload": "configs[2]: CLIP ViT-L/14 + LoRA r=16 (q,v) image tower, batch 512 per GPU, random-init weights",
               "images_per_s": L14_BATCH * world * args.steps / (ms_l / 1e3), "ms_per_step": ms_l / args.steps,
               "algorithmic_gflop_per_image": fl_l / 1e9, "tflops_per_gpu": tf_l,
               "frac_of_sustained_peak": tf_l / peaks["tf_sustained"], "frac_of_burst_peak": tf_l / peaks["tf_burst"],
               "kernel_ms_per_step": {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                                          "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1)}
                                      for k, v in prof_l.items()}}
        # configs[4] (seeker -> finder): the query side encodes 4096-caption batches with the ViT-L/14 text tower
        ids_q = synth_captions(QUERY_BATCH, seed=7)[0].to(dev, torch.int32)  # the same batch on every rank
        for _ in range(2):
            model_l.encode_texts(ids_q)
        ms_q = timed(lambda: model_l.encode_texts(ids_q), args.steps)
        l14["residual_stream"] = model_l.residual_dtype
        if model_l.residual_dtype != "float32":
            model_l.set_residual_dtype("float32")
            for _ in range(args.warmup):
                model_l.encode_images(pv_l)
            ms_l32 = timed(lambda: model_l.encode_images(pv_l), args.steps)
            tf_l32 = fl_l * L14_BATCH / (ms_l32 / args.steps * 1e-3) / 1e12
            l14["fp32_residual_stream"] = {"images_per_s": L14_BATCH * world * args.steps / (ms_l32 / 1e3),
                                           "ms_per_step": ms_l32 / args.steps, "tflops_per_gpu": tf_l32,
                                           "frac_of_sustained_peak": tf_l32 / peaks["tf_sustained"],
                                           "frac_of_burst_peak": tf_l32 / peaks["tf_burst"]}
            model_l.set_residual_dtype(l14["residual_stream"])
        l14["text_queries_per_s_per_gpu"] = QUERY_BATCH * args.steps / (ms_q / 1e3)
        l14["text_query_batch"] = QUERY_BATCH
        del pv_l
        if args.no_search:
            del model_l, ids_q
        torch.cuda.empty_cache()

    # ---- search: top-10 over the row-sharded 10M x 768 index ----------------------------
    search = None
    if not args.no_search:
        n_total = args.index_rows
        lo, hi = shard_bounds(n_total, rank, world)
        gi = torch.Generator(device=dev).manual_seed(4 + rank)
        shard = torch.empty((hi - lo, INDEX_DIM), dtype=torch.float32, device=dev)
        for s0 in range(0, hi - lo, 1_000_000):
            blk = torch.randn((min(1_000_000, hi - lo - s0), INDEX_DIM), generator=gi, device=dev)
            shard[s0:s0 + blk.shape[0]] = blk
        idx = TextSearchIndex(embeddings=shard, device=dev, distributed=world > 1, row_offset=lo,
                              total_rows=n_total, verbose=False)
        del shard
        gq = torch.Generator(device=dev).manual_seed(5)
        q = torch.randn((QUERY_BATCH, INDEX_DIM), generator=gq, device=dev)
        res = {}

        def sstep():
            res["s"], res["i"] = idx.search_batch(q, top_k=TOP_K)

        # streaming regime (HBM-bound): one 64-query tile scans the shard.  Measured FIRST and after
        # ~0.4 s of the same scan: after a tensor-bound phase the power-cap controller leaves the SM
        # clock near 0.9 GHz for ~100 ms, and the scan is L2-clock sensitive (profiles/r1_notes.md).
        q64 = q[:64].contiguous()
        for _ in range(128):  # fixed count: every rank must issue the same sequence of collectives
            idx.search_batch(q64, top_k=TOP_K)
        torch.cuda.synchronize()
        lib.clm_prof_enable(1)
        for _ in range(5):
            idx.search_batch(q64, top_k=TOP_K)
        p64 = _lib.prof_summary("search")
        lib.clm_prof_enable(0)
        n64 = max(1, p64["launches"])
        gbs = p64["bytes"] / (p64["ms"] * 1e-3) / 1e9
        for _ in range(args.warmup):
            sstep()
        ms_s = timed(sstep, args.steps)
        lib.clm_prof_enable(1)
        sstep()
        ps, pm = _lib.prof_summary("search"), _lib.prof_summary("merge")
        lib.clm_prof_enable(0)
        s_tf = ps["flops"] / (ps["ms"] * 1e-3) / 1e12
        # host-buffer end to end: pinned queries in, (score,id) out
        q_host = q.cpu().pin_memory()

        def s_e2e():
            s_, i_ = idx.search_batch(q_host.to(dev, non_blocking=True), top_k=TOP_K)
            s_.cpu(); i_.cpu()

        s_e2e()
        ms_se = timed(s_e2e, args.steps)
        search = {
            "metric": "top-10 queries/sec over 10M-row index", "value": QUERY_BATCH * args.steps / (ms_s / 1e3),
            "unit": "queries/s", "index": [n_total, INDEX_DIM], "rows_per_gpu": hi - lo, "query_batch": QUERY_BATCH,
            "k": TOP_K, "scaling": "strong (index rows fixed, sharded over GPUs)", "ms_per_batch": ms_s / args.steps,
            "exact": "bf16 tensor-core scan nominates every row within 2*eps (eps = 2^-8) of the k-th best bf16 score; fp32 "
                     "re-score of all of them; queries whose candidate list overflowed are redone with the exact fp32 scan",
            "e2e": {"value": QUERY_BATCH * args.steps / (ms_se / 1e3), "unit": "queries/s",
                    "h2d_bytes_per_step": q_host.numel() * 4, "d2h_bytes_per_step": QUERY_BATCH * TOP_K * 12},
            "roofline_q4096": {"bound": "tensor", "kernel": "search_kernel", "achieved": s_tf,
                               "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": s_tf / peaks["tf_sustained"],
                               "frac_of_burst": s_tf / peaks["tf_burst"],
                               "peak_source": f"{peaks['source']} bf16_tflops_sustained (scan runs under the power cap)",
                               "scan_ms": ps["ms"], "merge_ms": pm["ms"]},
            "roofline_q64": {"bound": "hbm", "kernel": "search_kernel", "achieved": gbs, "peak": peaks["hbm_gbs"],
                             "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "scan_ms": p64["ms"] / n64,
                             "bytes_per_scan": p64["bytes"] / n64, "queries": 64},
        }
        # configs[4]: top-50 over the same index (the seeker path's k) ...
        for _ in range(2):
            idx.search_batch(q, top_k=50)
        ms_50 = timed(lambda: idx.search_batch(q, top_k=50), args.steps)
        search["top50"] = {"value": QUERY_BATCH * args.steps / (ms_50 / 1e3), "unit": "queries/s", "k": 50,
                           "ms_per_batch": ms_50 / args.steps}
        # ... and the whole seeker -> finder step, timed as one: every rank encodes ITS block of the 4096-caption
        # query batch with the ViT-L/14+LoRA text tower (data parallel), one all_gather rebuilds the [Q, 768]
        # query matrix, then the row-sharded top-50 search (local scan + all_gather of Q x 50 pairs + merge)
        if l14 is not None:
            from clip_lora_match_b200.src.embedding.search import allgather_rows

            qlo, qhi = shard_bounds(QUERY_BATCH, rank, world)
            ids_mine = ids_q[qlo:qhi].contiguous()

            def seeker_step():
                emb = model_l.encode_texts(ids_mine)
                if world > 1:
                    emb = allgather_rows(emb, QUERY_BATCH)
                res["s50"], res["i50"] = idx.search_batch(emb, top_k=50)

            for _ in range(2):
                seeker_step()
            ms_sk = timed(seeker_step, args.steps)
            search["seeker_config5"] = {
                "value": QUERY_BATCH * args.steps / (ms_sk / 1e3), "unit": "queries/s", "ms_per_batch": ms_sk / args.steps,
                "what": "configs[4]: ViT-L/14+LoRA text encode of a 4096-query batch (split over the ranks, one "
                        "all_gather of the embeddings) + top-50 over the row-sharded 10M x 768 index, one timed step"}
            del model_l, ids_q
        del idx
        torch.cuda.empty_cache()

    # ---- LoRA training step (reference scripts/train_lora.py) -------------------
    # The reference's own training configuration (config/lora_config.yaml: ViT-B/32, r = 8 on q,k,v,out, batch 8,
    # temperature 0.07, AdamW + clip) and a batch of 256: forward + backward + clip + AdamW per step through
    # LoraTrainer.step (CUDA-graph replay), device-resident inputs, and end to end from pinned host inputs with
    # the loss read back.  N > 1: data-parallel (an extension: the reference's script is one process) -- every rank
    # steps on its own batch and the flat fp32 gradient buffer is all-reduced once per step over NCCL (weak scaling).
    train = None
    if not args.no_train:
        from clip_lora_match_b200.models.lora_trainer import LoraTrainer

        arch_t = CM.arch_from_name(TRAIN_ARCH)
        sd_t = CM.random_init_state_dict(arch_t, seed=0)
        train = {"workload": "reference config/lora_config.yaml: CLIP ViT-B/32 + LoRA r=8 (q,k,v,out), symmetric "
                             "InfoNCE T=0.07, clip 1.0 + AdamW, random-init weights, synthetic pairs; LoRA dropout not applied",
                 "batches": {}}
        for tb in TRAIN_BATCHES:
            model_t = CM.B200ClipModel(arch_t, sd_t, device=dev)
            model_t.set_lora(init_lora_adapter(model_t.linear_dims(), LoraConfig(r=8, lora_alpha=16, target_modules=TRAIN_TARGETS),
                                               seed=1, base_model_name=TRAIN_ARCH))
            tr = LoraTrainer(model_t, lr=1e-4, weight_decay=0.01, max_grad_norm=1.0, temperature=0.07,
                             distributed=world > 1)
            gt = torch.Generator(device=dev).manual_seed(8 + rank)
            pv_t = torch.randn((tb, 3, 224, 224), generator=gt, device=dev)
            ids_t = synth_captions(tb, seed=9 + rank)[0].to(dev, torch.int32)
            pv_h, ids_h = pv_t.cpu().pin_memory(), ids_t.cpu().pin_memory()
            for _ in range(max(args.warmup, 3)):
                tr.step(pv_t, ids_t)
            n0 = lib.clm_launch_count()
            ms_t = timed(lambda: tr.step(pv_t, ids_t), args.steps)
            launches_t = (lib.clm_launch_count() - n0) // args.steps
            loss_last = []

            def t_e2e():
                loss_last.append(tr.step(pv_h.to(dev, non_blocking=True), ids_h.to(dev, non_blocking=True)).item())

            ms_te = timed(t_e2e, args.steps)
            lib.clm_prof_enable(1)
            tr.step(pv_t, ids_t)
            prof_t = {kn: _lib.prof_summary(kn) for kn in ("gemm", "attention", "elementwise")}
            lib.clm_prof_enable(0)
            train["batches"][str(tb)] = {
                "pairs_per_s": tb * world * args.steps / (ms_t / 1e3), "ms_per_step": ms_t / args.steps,
                "e2e_pairs_per_s": tb * world * args.steps / (ms_te / 1e3), "e2e_ms_per_step": ms_te / args.steps,
                "batch_per_gpu": tb, "n_gpus": world, "scaling": "weak",
                "exchange_bytes_per_step": (tr.grad.numel() * 4 + 4) if world > 1 else 0,
                "h2d_bytes_per_step": pv_h.numel() * 4 + ids_h.numel() * 4, "d2h_bytes_per_step": 4,
                "launches_per_step": int(launches_t), "trainable_parameters": tr.num_trainable_parameters(),
                "loss_first_last": [loss_last[0], loss_last[-1]],
                "kernel_ms_eager_step": {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                                             "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1)}
                                         for k, v in prof_t.items()}}
            del tr, model_t
            torch.cuda.empty_cache()
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_train_rate()
            train["cpu_baseline"] = {"value": v, "unit": "image-caption pairs/s", "cores": cores, "kind": "port",
                                     "sample": sample}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_encode_rate()
        cpu = {"value": v, "unit": "image-caption pairs/s", "cores": cores, "kind": "port", "sample": sample}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        line = {
            "metric": "CLIP+LoRA images/sec", "value": value, "unit": "image-caption pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "detail": detail, "vit_l14": l14, "search": search, "train": train,
        }
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
