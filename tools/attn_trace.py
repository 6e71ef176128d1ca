"""Timeline of the attention kernel's hand-offs (CTA 0, first tiles).  Builds a -DCLM_ATTN_TRACE copy of
the library into tools/_trace/ and prints, per tile, when each warp passed each hand-off point.
Usage: python tools/attn_trace.py [--T 197] [--H 12] [--B 64]"""
import argparse, ctypes as C, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "clip_lora_match_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "_trace", "libclm_trace.so")

def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [os.path.join(CSRC, "clm_api.cu"), os.environ.get("CLM_TRACE_SRC", os.path.join(CSRC, "clm_attention.cu"))]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DCLM_ATTN_TRACE", "-Xcompiler", "-fPIC",
           "-shared", "-cudart", "static", "-o", OUT] + srcs
    subprocess.run(cmd, check=True)

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--T", type=int, default=197); ap.add_argument("--H", type=int, default=12)
    ap.add_argument("--B", type=int, default=256); ap.add_argument("--causal", type=int, default=0); ap.add_argument("--build-only", action="store_true")
    a = ap.parse_args()
    if not os.path.exists(OUT) or a.build_only: build()
    if a.build_only: return
    lib = C.CDLL(OUT)
    dev = torch.device("cuda")
    NW, NT, NE = 24, 12, 8
    buf = torch.zeros(NW * NT * NE, dtype=torch.int64, device=dev)
    qkv = torch.randn((a.B * a.T, 3 * a.H * 64), device=dev).bfloat16()
    out = torch.empty((a.B * a.T, a.H * 64), dtype=torch.bfloat16, device=dev)
    lib.clm_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.clm_attention_set_trace.argtypes = [C.c_void_p]
    for _ in range(3):
        assert lib.clm_attention(qkv.data_ptr(), out.data_ptr(), a.B, a.T, a.H, a.causal, None) == 0
    torch.cuda.synchronize()
    assert lib.clm_attention_set_trace(buf.data_ptr()) == 0
    assert lib.clm_attention(qkv.data_ptr(), out.data_ptr(), a.B, a.T, a.H, a.causal, None) == 0
    torch.cuda.synchronize()
    tr = buf.cpu().view(NW, NT, NE)
    t0 = int(tr[tr > 0].min())
    names = ["wait_S", "got_S", "pass1_done", "max_xchg", "P_published", "got_O", "O_in_regs", "stage_ok"]
    print(f"# T={a.T} H={a.H} B={a.B}: clocks relative to the first stamp; MMA warp (1): ev0 = S issued, ev1 = PV issued")
    env = os.environ.get("CLM_ATTN_SPLIT", "1")
    split = not a.causal and ((a.T == 257 and env != "0") or ((a.T + 15) // 16 * 16 == 208 and a.T > 192 and env == "2"))
    mma_warps = (22, 23) if split else (1,)
    tails = (0, 1, 2, 3) if split else (18, 19)
    tma = 20 if split else 0
    softs = range(4, 20) if split else range(2, 18)
    s0 = 4 if split else 2
    for t in range(NT):
        for w in mma_warps:
            if tr[w, t, 0] or tr[w, t, 1]:
                print(f"tile {t}: MMA warp {w} S_issue={int(tr[w,t,0])-t0 if tr[w,t,0] else None} PV_issue={int(tr[w,t,1])-t0 if tr[w,t,1] else None}")
        if tr[tma, t, 1]:
            print(f"   TMA warp, item {t}: wants_stage={int(tr[tma,t,0])-t0} stage_free={int(tr[tma,t,1])-t0}")
        for w in tails:
            ev = [int(x) - t0 if x else None for x in tr[w, t].tolist()]
            if any(e is not None for e in ev):
                print(f"   extra-token warp {w}, item {t}: wait_stage={ev[0]} got_stage={ev[1]} scores_done={ev[2]} row_done={ev[3]}")
        for w in softs:
            ev = [int(x) - t0 if x else None for x in tr[w, t].tolist()]
            if any(e is not None for e in ev):
                role = f"grp {((w-s0)>>2)&1} half {(w-s0)>>3} q {w&3}"
                print(f"   warp {w:2d} ({role}): " + " ".join(f"{n}={e}" for n, e in zip(names, ev)))

if __name__ == "__main__":
    main()
