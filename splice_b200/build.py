"""In-tree build of the splice_b200 CUDA extension (plain C-ABI shared library, no torch headers).

`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` on every `csrc/*.cu`, linked into
`splice_b200/libsplice_b200.so`. Objects are cached under `splice_b200/csrc/_obj/` keyed by source mtime, so
repeated calls are cheap. nvcc cross-compiles without a GPU, so this runs on the CPU-only build box too.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OBJ_DIR = CSRC / "_obj"
LIB_PATH = PKG_DIR / "libsplice_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-I", str(PKG_DIR.parent / "include"),
    "-I", str(CSRC),
]
# SPLICE_B200_CROSSCHECK=1: also build the cross-check kernels (mma.sync attention, one-tile-per-CTA GEMM, the 96-wide and
# thread-block-cluster GEMM shapes) that A/B tools select through SPLICE_B200_ATTN / impl / bn_hint. Not in the product library.
if os.environ.get("SPLICE_B200_CROSSCHECK", "0") == "1":
    NVCC_FLAGS.append("-DSPLICE_B200_CROSSCHECK")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stamp(src: Path, headers: list[Path]) -> str:
    h = hashlib.sha1()
    for p in [src, *headers]:
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


STAMP_PATH = PKG_DIR / "libsplice_b200.so.stamp"


def tree_stamp() -> str:
    """Hash of everything the library is built from (sources, headers, flags): written next to the .so by build(), compared
    by _lib at import so that a stale library is never loaded silently."""
    h = hashlib.sha1()
    files = sorted(CSRC.glob("*.cu")) + sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((PKG_DIR.parent / "include").glob("*.h")))
    for p in files:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(f for f in NVCC_FLAGS if not os.path.isabs(f)).encode())   # include paths differ between machines
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False) -> Path:
    """Compile every .cu for sm_100a and link the shared library. Returns the .so path."""
    OBJ_DIR.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((PKG_DIR.parent / "include").glob("*.h")))
    if not sources:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    nvcc = _nvcc()
    jobs = []
    objs = []
    for src in sources:
        obj = OBJ_DIR / (src.stem + ".o")
        stamp_file = OBJ_DIR / (src.stem + ".stamp")
        stamp = _stamp(src, headers)
        objs.append(obj)
        if force or not obj.exists() or not stamp_file.exists() or stamp_file.read_text() != stamp:
            jobs.append((src, obj, stamp_file, stamp))

    def compile_one(job):
        src, obj, stamp_file, stamp = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        stamp_file.write_text(stamp)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not LIB_PATH.exists() or force:
        cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    STAMP_PATH.write_text(tree_stamp())
    return LIB_PATH


if __name__ == "__main__":
    p = build(verbose=True, force="--force" in sys.argv)
    print(p)
