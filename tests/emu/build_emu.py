"""Builds tests/emu/_build/libgenx_emu.so: csrc/generator_x.cu compiled by g++ as plain C++ against cuda_emu.h (TEST
INFRASTRUCTURE: the same kernel bodies and host orchestration as the product library, executed by CPU fibers). Nothing in
splice_b200/ imports this; tests/test_genx_emu.py does."""
from __future__ import annotations

import hashlib
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
CSRC = ROOT / "splice_b200" / "csrc"
OUT = HERE / "_build"
LIB = OUT / "libgenx_emu.so"
SOURCES = [CSRC / "generator_x.cu", CSRC / "generator_x.h", CSRC / "gen_dev.cuh", CSRC / "gen_kernels.cuh", HERE / "cuda_emu.h",
           ROOT / "include" / "splice_b200.h"]
FLAGS = ["-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DSPLICE_EMU", "-Wno-unknown-pragmas", "-fvisibility=hidden",
         "-I", str(HERE), "-I", str(CSRC), "-I", str(ROOT / "include")]


def build() -> Path:
    OUT.mkdir(exist_ok=True)
    h = hashlib.sha1()
    for p in SOURCES:
        h.update(p.read_bytes())
    h.update(" ".join(FLAGS).encode())
    stamp = OUT / "stamp"
    if LIB.exists() and stamp.exists() and stamp.read_text() == h.hexdigest():
        return LIB
    cmd = ["g++", *FLAGS, "-x", "c++", str(CSRC / "generator_x.cu"), "-o", str(LIB)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"emulation build failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(h.hexdigest())
    return LIB


if __name__ == "__main__":
    print(build())
