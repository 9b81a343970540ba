"""Print the lowered k_tile3 program (arm sequence) of every fused pass of a workload -- host code only, no GPU.

    python tools/dump_tile3.py qft 30        python tools/dump_tile3.py layered 30 [-v]
"""
import ctypes as C
import collections
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

import spinoza_b200 as sb
from spinoza_b200 import QuantumCircuit, QuantumRegister, workloads

EMU = ROOT / "tests" / "emu"
lib = EMU / "_build" / "libtile_emu.so"
(EMU / "_build").mkdir(exist_ok=True)
subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-pthread", "-I/usr/local/cuda/include",
                "-include", str(EMU / "cuda_cpu_shim.h"), "-x", "c++", str(EMU / "tile_emu.cpp"), "-o", str(lib)], check=True, cwd=ROOT)
h = C.CDLL(str(lib))
h.emu_tile3_dump.restype = C.c_int
h.emu_tile3_dump.argtypes = [C.c_int, C.c_char_p, C.c_longlong, C.c_void_p, C.c_int, C.POINTER(C.c_int)]

what = sys.argv[1] if len(sys.argv) > 1 else "qft"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
verbose = "-v" in sys.argv
qc = QuantumCircuit(QuantumRegister(n), fuse=True) if hasattr(sb, "QuantumRegister") else None
if what == "qft":
    qc.qft()
else:
    workloads.random_layered_circuit(qc, depth=20, seed=42)
plan, n_pass = qc.plan()
arr, cnt = qc._encode()
NAMES = {0: "END", 1: "LAYOUT", 2: "ACC", 3: "ACCG", 4: "OTHER"}
GV = ["H", "RX", "RY", "HS.g", "RX.g", "RY.g", "X.g", "Y.g"]


def name(op):
    if op in NAMES:
        return NAMES[op]
    if 5 <= op < 9:
        return f"PRE{op - 5}"
    return f"{GV[(op - 9) // 4]}{(op - 9) % 4}"


total = collections.Counter()
for p in range(n_pass):
    buf = (C.c_char * (8 << 20))()
    used = C.c_int64()
    sb._check(sb._lib.spz_debug_compile_pass(n, arr, cnt, qc._flags(), p, buf, len(buf), C.byref(used)))
    blob = bytes(buf[: used.value])
    if int(np.frombuffer(blob, dtype="<i4", count=1)[0]) == 1:
        print(f"pass {p}: single op (direct kernel)")
        continue
    ins = np.zeros((4096, 32), dtype=np.uint8)  # Ins3 is 32 bytes: operand words, then the instruction's own two scalars
    info = (C.c_int * 4)()
    k = h.emu_tile3_dump(n, blob, len(blob), ins.ctypes.data, 4096, info)
    ops = [name(int(o)) for o in ins[:k, 0]]
    c = collections.Counter(o.rstrip("0123") for o in ops)
    total.update(c)
    hdr = np.frombuffer(blob, dtype="<i4", count=16)
    print(f"pass {p}: L={hdr[2]} high={list(hdr[4:4 + hdr[3]])} ins={k} groups={info[2]} terms={info[3]} " + " ".join(f"{a}:{b}" for a, b in sorted(c.items())))
    if verbose:
        print("   " + " ".join(ops))
    if "-d" in sys.argv:
        rec = ins[:k].copy().view(np.dtype([("op", "u1"), ("kind", "u1"), ("rpos", "u1"), ("flags", "u1"), ("km", "<u2"), ("thr", "<u2"),
                                            ("a", "<u4"), ("b", "<u4"), ("s0", "<f8"), ("s1", "<f8")])).ravel()
        for r in rec:
            print(f"      {name(int(r['op'])):8s} cls/r {r['rpos']} flags {int(r['flags']):08b} km {int(r['km']):016b} thr {int(r['thr']):08b} a {r['a']:#x} b {r['b']}")
print("total", dict(total))
