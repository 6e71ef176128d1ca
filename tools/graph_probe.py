"""Does launch overhead matter at batch 1024?  Times the configs[1] step launched eagerly and replayed from a
CUDA graph captured over the same library calls.  Usage: python tools/graph_probe.py [--batch 1024] [--arch ...]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_lora_match_b200.models import clip_model as CM
from clip_lora_match_b200.models.lora_adapter import LoraConfig, init_lora_adapter

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="openai/clip-vit-base-patch16"); ap.add_argument("--batch", type=int, nargs="+", default=[1024, 1])
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda")
    arch = CM.arch_from_name(a.arch)
    model = CM.B200ClipModel(arch, CM.random_init_state_dict(arch, 0), device=dev)
    model.set_lora(init_lora_adapter(model.linear_dims(), LoraConfig(r=16, lora_alpha=32), seed=1, init_b_std=0.02))
    for B in a.batch:
        pv = torch.randn((B, 3, 224, 224), device=dev)
        ids = torch.randint(0, 49405, (B, 77), device=dev, dtype=torch.int32); ids[:, 0] = 49406; ids[:, 40:] = 49407
        def step():
            return model.encode_images(pv), model.encode_texts(ids)
        for _ in range(3): ref = step()
        torch.cuda.synchronize()
        steps = a.steps if B > 16 else 50
        def timeit(fn):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps
        import time
        t0 = time.perf_counter(); 
        for _ in range(steps): step()
        host_ms = (time.perf_counter() - t0) / steps * 1e3   # host time to ENQUEUE one step (no sync inside)
        torch.cuda.synchronize()
        eager = timeit(step)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = step()
        g.replay(); torch.cuda.synchronize()
        ok = all(torch.equal(x, y) for x, y in zip(out, ref))
        graph = timeit(g.replay)
        eager2 = timeit(step)
        print(json.dumps({"arch": a.arch, "batch": B, "eager_ms": round(eager, 4), "graph_ms": round(graph, 4), "eager_again_ms": round(eager2, 4),
                          "host_enqueue_ms": round(host_ms, 4), "graph_equals_eager": ok}), flush=True)

if __name__ == "__main__":
    main()
