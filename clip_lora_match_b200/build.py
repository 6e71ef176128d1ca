"""In-tree build of libclm_b200.so (sm_100a only) with plain nvcc.

`python -m clip_lora_match_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles
without a GPU; the .so is git-ignored but travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = CSRC / "libclm_b200.so"
SOURCES = [
    "clm_api.cu",
    "clm_elementwise.cu",
    "clm_preprocess.cu",
    "clm_gemm.cu",
    "clm_attention.cu",
    "clm_search.cu",
    "clm_tower.cu",
    "clm_train.cu",
    "clm_attention_bwd.cu",
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--threads", "2",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the sm_100a extension cannot be built")


def _stamp() -> str:
    h = hashlib.sha256()
    for name in sorted(p.name for p in CSRC.glob("*.cu*")):
        h.update(name.encode())
        h.update((CSRC / name).read_bytes())
    h.update((CSRC.parent.parent / "include" / "clm_b200.h").read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> Path:
    """Compile every CUDA source for sm_100a and link libclm_b200.so next to the sources."""
    stamp_file = CSRC / ".build_stamp"
    stamp = _stamp()
    if not force and LIB_PATH.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB_PATH
    nvcc = _nvcc()
    objdir = CSRC / "build"
    objdir.mkdir(exist_ok=True)

    def compile_one(src: str) -> Path:
        obj = objdir / (src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH),
            *map(str, objs), "-cudart", "static"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp_file.write_text(stamp)
    if verbose:
        print(f"[clm_b200.build] built {LIB_PATH} ({LIB_PATH.stat().st_size >> 10} KiB)")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
