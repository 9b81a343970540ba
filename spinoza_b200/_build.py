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
    deps.append(PKG / "ptx_jump_table.py")
    deps.append(PKG / "_build.py")
    for d in deps:
        h.update(d.name.encode())
        h.update(d.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("SPZ_NO_JUMP_TABLE", "").encode())
    return h.hexdigest()


def _stale() -> bool:
    # content hash, not mtimes: the snapshot that travels to the GPU box does not preserve mtimes
    if not LIB.exists() or not STAMP.exists():
        return True
    return STAMP.read_text().strip() != _source_hash()


PTX_PATCHED = ["kernels_tile3.cu"]  # sources whose dense switch becomes a jump table in the PTX (see ptx_jump_table.py)
JT_REPORT = LIBDIR / ".jump_table"


def _compile_with_jump_table(nvcc: str, flags, src: Path, obj: Path, env, ccbin) -> str:
    """nvcc's own steps (taken from --dryrun), with the PTX rewritten between cicc and ptxas.  Returns a one-line report;
    raises on any failure (the caller then compiles the source the ordinary way)."""
    import re
    import shlex
    from . import ptx_jump_table
    keep = obj.parent / (src.name + ".keep")
    shutil.rmtree(keep, ignore_errors=True)
    keep.mkdir(parents=True)
    cmd = [nvcc, *flags, "--dryrun", "--keep", "--keep-dir", str(keep), "-c", str(src), "-o", str(obj)]
    if ccbin:
        cmd += ["-ccbin", ccbin]
    dry = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if dry.returncode != 0:
        raise RuntimeError(dry.stderr)
    step_env = dict(env)
    report = None
    for line in (dry.stderr + dry.stdout).splitlines():
        if not line.startswith("#$ "):
            continue
        body = line[3:].strip()
        m = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)=(\S*)$", body)
        if m:
            step_env[m.group(1)] = m.group(2).strip('"')
            continue
        if re.match(r"^[A-Za-z_][A-Za-z0-9_]*=", body):
            continue  # (multi-word settings such as LIBRARIES: not used by the steps)
        if body.startswith("rm "):
            continue
        if re.match(r'^"?ptxas"?\s', body) or " ptxas " in body.split("-o")[0][:20]:
            ptx = next(Path(a) for a in shlex.split(body) if a.endswith(".ptx"))
            text, n = ptx_jump_table.patch(ptx.read_text())
            if n != 1:
                raise RuntimeError(f"{n} switches rewritten in {ptx.name}, expected 1")
            ptx.write_text(text)
            report = f"{src.name}: switch -> brx.idx, {ptx_jump_table.arms_reached(text)} targets"
        r = subprocess.run(["bash", "-c", body], capture_output=True, text=True, env=step_env, cwd=str(keep))
        if r.returncode != 0:
            raise RuntimeError(f"step failed: {body[:200]}\n{r.stdout}{r.stderr}")
    if report is None or not obj.exists():
        raise RuntimeError("no ptxas step found in nvcc --dryrun")
    cub = next(keep.glob("*.cubin"), None)
    if cub is not None:
        sass = subprocess.run(["cuobjdump", "-sass", str(cub)], capture_output=True, text=True).stdout
        report += f"; SASS: {sass.count(' BRX ')} BRX"
    shutil.rmtree(keep, ignore_errors=True)
    return report


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
    jt_lines = []
    for src in SOURCES:
        if not (CSRC / src).exists():
            continue
        obj = objdir / (src + ".o")
        objs.append(str(obj))
        if src in PTX_PATCHED and os.environ.get("SPZ_NO_JUMP_TABLE") != "1":
            try:
                jt_lines.append(_compile_with_jump_table(nvcc, [*NVCC_FLAGS, "-Xptxas", "-warn-spills"], CSRC / src, obj, env, ccbin))
                continue
            except Exception as e:  # the ordinary build is always possible: same results, compare-tree dispatch
                jt_lines.append(f"{src}: jump table NOT applied ({str(e).splitlines()[0][:160] if str(e) else type(e).__name__})")
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
    JT_REPORT.write_text("\n".join(jt_lines) + "\n")
    if verbose:
        print("\n".join(jt_lines))
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
