"""Build a variant of libclm_b200.so with extra -D defines, for same-run A/B measurements on the GPU box
(run a tool with CLM_LIB_PATH=<variant>).  Usage: python tools/build_variant.py OUT.so -DNAME=VALUE [...]"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lora_match_b200.build import CSRC, NVCC_FLAGS, SOURCES, _nvcc

def main():
    out, defs = sys.argv[1], sys.argv[2:]
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, *defs, "-shared", "-cudart", "static", "-o", out] + [str(CSRC / s) for s in SOURCES]
    subprocess.run(cmd, check=True)
    print("built", out)

if __name__ == "__main__":
    main()
