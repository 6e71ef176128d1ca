"""Debug: batch-1 text encode vs oracle for several lengths (tiny-test)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import clip_oracle as O
from clip_lora_match_b200.models import clip_model as CM

dev = torch.device("cuda:0")
model = O.build_model("tiny-test", seed=0)
arch = CM.arch_from_hf_config(O.hf_config("tiny-test"), "tiny-test")
gpu = CM.B200ClipModel(arch, O.base_state_dict(model), device=dev)
cos = torch.nn.functional.cosine_similarity
g = torch.Generator().manual_seed(0)
for B in (1, 2, 3):
    for L in (5, 10, 33, 64, 65, 77):
        ids = torch.randint(1, 49000, (B, L), generator=g)
        ids[:, 0] = 49406; ids[:, -1] = 49407
        ref = O.encode_texts(model, ids, None)
        got = gpu.encode_texts(ids).cpu()
        ids77 = torch.full((B, 77), 49407, dtype=torch.long); ids77[:, :L] = ids
        ref77 = O.encode_texts(model, ids77, None)
        print(f"B={B} L={L}: cos(gpu, oracle unpadded)={cos(got, ref).min():.5f} cos(gpu, oracle padded)={cos(got, ref77).min():.5f} cos(oracle pad, unpad)={cos(ref, ref77).min():.5f}", flush=True)
proc = CM.ClmProcessor("openai/clip-vit-base-patch32")
txt = "tas pink kanken, ditemukan di lab iot"
ids = proc(text=[txt])["input_ids"]
print("ids", ids.tolist())
t = CM.encode_text(txt, gpu, proc, dev)
print("cos api vs oracle", cos(t, O.encode_texts(model, ids, None)[0], dim=0).item())
