"""Compile the C part of the oracle (TEST INFRASTRUCTURE): gcc -> oracle/_build/libclm_oracle.so."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_build" / "libclm_oracle.so"
SOURCES = ["pil_resample.c"]


def build(force: bool = False) -> Path:
    srcs = [HERE / s for s in SOURCES]
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    # -ffp-contract=off: Pillow's x86-64 build has no fused multiply-add; keep every rounding it has
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(OUT), *map(str, srcs), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{r.stdout}\n{r.stderr}")
    return OUT


def load() -> C.CDLL:
    lib = C.CDLL(str(build()))
    lib.clm_oracle_resize_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.clm_oracle_resized_shape.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.clm_oracle_clip_preprocess.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p]
    return lib


if __name__ == "__main__":
    print(build(force=True))
