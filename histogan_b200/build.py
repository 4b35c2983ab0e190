"""Build libhistogan_b200.so (sm_100a only) in-tree with nvcc.

    python -m histogan_b200.build [--force] [--verbose]

The library is plain C-ABI (include/histogan_b200.h); it links the static CUDA
runtime, so it only needs the NVIDIA driver at run time.  nvcc cross-compiles
without a GPU, so this also runs on the CPU-only build box.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIBDIR = HERE / "lib"
LIB = LIBDIR / "libhistogan_b200.so"
OBJDIR = HERE / "build"

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libhistogan_b200.so")


def sources():
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) +
                    [HERE.parent / "include" / "histogan_b200.h", Path(__file__)]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = LIBDIR / ".fingerprint"
    fp = _fingerprint()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB
    nvcc = _nvcc()
    OBJDIR.mkdir(exist_ok=True)
    LIBDIR.mkdir(exist_ok=True)

    # headers every translation unit depends on; an object is rebuilt only when its own source,
    # a header or the flags changed
    hdr = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "histogan_b200.h"]):
        hdr.update(f.read_bytes())
    hdr.update(" ".join(ARCH_FLAGS + NVCC_FLAGS).encode())

    def compile_one(src: Path) -> Path:
        obj = OBJDIR / (src.stem + ".o")
        ostamp = OBJDIR / (src.stem + ".sha")
        key = hashlib.sha256(hdr.digest() + src.read_bytes()).hexdigest()
        if not force and obj.exists() and ostamp.exists() and ostamp.read_text() == key:
            return obj
        cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJDIR / (src.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        ostamp.write_text(key)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(fp)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
