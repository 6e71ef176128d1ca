"""Debug: q64 scan time vs SM clock after a heavy q4096 phase."""
import os, sys, json, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_lora_match_b200 import _lib, kernels as K
dev = torch.device("cuda"); lib = _lib.load()
N, D = 10_000_000, 768
g = torch.Generator(device=dev).manual_seed(4)
e = torch.randn((N, D), generator=g, device=dev); e = e / e.norm(dim=-1, keepdim=True)
eb = e.bfloat16()
q = torch.randn((4096, D), generator=g, device=dev); q = q / q.norm(dim=-1, keepdim=True)
qb = q.bfloat16(); q64, qb64 = q[:64].contiguous(), qb[:64].contiguous()
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu", "--format=csv,noheader,nounits", "-lms", "50"],
                        stdout=subprocess.PIPE, text=True)
t00 = time.time()
def rd():
    for line in proc.stdout: rows.append((time.time() - t00, line.strip()))
threading.Thread(target=rd, daemon=True).start()
def phase(name, fn, seconds):
    t0 = time.time(); out = []
    while time.time() - t0 < seconds:
        lib.clm_prof_enable(1); fn(); ps = _lib.prof_summary("search"); lib.clm_prof_enable(0)
        out.append((round(time.time() - t00, 3), round(ps["ms"], 3)))
    print(name, out[:3], "...", out[-3:], "n=", len(out), flush=True)
    return out
phase("q64 cold", lambda: K.search_topk(q64, qb64, eb, e, 10), 1.0)
phase("q4096", lambda: K.search_topk(q, qb, eb, e, 10), 2.0)
o = phase("q64 after heavy", lambda: K.search_topk(q64, qb64, eb, e, 10), 3.0)
print("q64 trace:", o[::10])
time.sleep(1.0)
phase("q64 after 1s idle", lambda: K.search_topk(q64, qb64, eb, e, 10), 1.0)
proc.terminate()
print("clock samples (t, sm MHz, mem MHz, W, C):")
for r in rows[::4]: print(" ", round(r[0], 2), r[1])
