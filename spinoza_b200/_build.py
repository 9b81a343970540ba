"""Builds libspinoza_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

Usage: ``python -m spinoza_b200._build [--force]`` or ``spinoza_b200._build.build()``.
The .so lands in spinoza_b200/lib/ (git-ignored, but it travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libspinoza_b200.so"
SOURCES = ["abi.cu", "kernels_direct.cu", "kernels_reduce.cu", "kernels_tile.cu", "kernels_tile3.cu", "dist.cu", "rendezvous.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libspinoza_b200.so")


STAMP = LIBDIR / ".source_hash"


def _source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    deps = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inl")))
    deps.append(PKG.parent / "include" / "spinoza_b200.h")
    for d in deps:
        h.update(d.name.encode())
        h.update(d.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _stale() -> bool:
    # content hash, not mtimes: the snapshot that travels to the GPU box does not preserve mtimes
    if not LIB.exists() or not STAMP.exists():
        return True
    return STAMP.read_text().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB
    LIBDIR.mkdir(exist_ok=True)
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    # the image exports CC=/opt/gcc/bin/gcc, whose OpenMP spec is missing; nvcc only needs a host g++
    ccbin = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else None
    procs = []
    objs = []
    for src in SOURCES:
        if not (CSRC / src).exists():
            continue
        obj = objdir / (src + ".o")
        objs.append(str(obj))
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", str(CSRC / src), "-o", str(obj)]
        if ccbin:
            cmd += ["-ccbin", ccbin]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    if ccbin:
        cmd += ["-ccbin", ccbin]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    STAMP.write_text(_source_hash())
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
