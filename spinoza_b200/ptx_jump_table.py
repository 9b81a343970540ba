"""Turns the compare-and-branch tree nvcc emits for a dense `switch` into one indirect branch (PTX `brx.idx`, SASS `BRX`).

nvcc 12.9 lowers every C++ `switch` to a balanced tree of `setp` / `bra` pairs -- it never emits a jump table (checked on a
41-case toy: no `brx.idx` in the PTX, no `BRX` in the SASS).  For k_tile3's interpreter that is six dependent compare + branch
levels per interpreted instruction, each taken branch a fresh instruction fetch: measured with ncu, the dispatch of an RY
costs as many cycles as its 48 DFMAs.  There is no way to write an indirect branch in CUDA C++ (no computed goto, no
`asm goto`; labels of separate asm statements are not visible to each other), but ptxas accepts `brx.idx` with a
`.branchtargets` table.  So the build compiles the kernel to PTX, rewrites the tree that follows a marker comment
(`asm volatile("// SPZ_JUMP_TABLE <n>")` placed right before the `switch`) and hands the PTX to ptxas:

    marker ... setp.gt.s16 %p, %rs2, 20; @%p bra A; setp ... (the tree, possibly over several basic blocks)
 -> marker ... brx.idx %r_idx, TABLE;  TABLE: .branchtargets L_0, L_1, ..., L_{n-1};

The targets are found by *evaluating* the tree for every selector value 0 .. n-1 (an interpreter for the handful of PTX
instructions a switch tree consists of: setp.{eq,ne,lt,le,gt,ge} on the selector register against immediates, predicated
and uniform branches, labels); anything else ends the walk: that instruction starts the case's body.  The dead tree is left
for ptxas to delete.  If the marker or the expected shape is not found the PTX is returned unchanged and the caller builds the
unpatched kernel -- same results, slower dispatch (build() reports which one was built).
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Tuple

_SETP = re.compile(r"^\s*setp\.(eq|ne|lt|le|gt|ge|lo|ls|hi|hs)\.([su])(16|32)\s+(%p\d+),\s*(%\w+),\s*(-?\d+);")
_PBRA = re.compile(r"^\s*@(!?)(%p\d+)\s+bra(?:\.uni)?\s+(\$?[\w$]+);")
_BRA = re.compile(r"^\s*bra(?:\.uni)?\s+(\$?[\w$]+);")
_LABEL = re.compile(r"^(\$?[\w$]+):")
_SKIP = re.compile(r"^\s*(\.loc\b.*|//.*|)$")

_CMP = {"eq": lambda a, b: a == b, "ne": lambda a, b: a != b, "lt": lambda a, b: a < b, "le": lambda a, b: a <= b,
        "gt": lambda a, b: a > b, "ge": lambda a, b: a >= b, "lo": lambda a, b: a < b, "ls": lambda a, b: a <= b,
        "hi": lambda a, b: a > b, "hs": lambda a, b: a >= b}


class PatchError(Exception):
    pass


def _walk(lines: List[str], labels: Dict[str, int], start: int, sel: str, value: int) -> int:
    """Follow the switch tree from line `start` for selector value `value`; returns the index of the first line that is not part
    of the tree (the case body)."""
    preds: Dict[str, bool] = {}
    i = start
    for _ in range(10000):
        ln = lines[i]
        if _SKIP.match(ln) or _LABEL.match(ln):
            i += 1
            continue
        m = _SETP.match(ln)
        if m and m.group(5) == sel:
            preds[m.group(4)] = _CMP[m.group(1)](value, int(m.group(6)))
            i += 1
            continue
        m = _PBRA.match(ln)
        if m and m.group(2) in preds:
            taken = preds[m.group(2)] != (m.group(1) == "!")
            i = labels[m.group(3)] if taken else i + 1
            continue
        m = _BRA.match(ln)
        if m:
            # a uniform branch is part of the tree only while we have not left it: follow it, the body starts at its target
            i = labels[m.group(1)]
            continue
        return i
    raise PatchError("switch tree does not terminate")


def patch(ptx: str, marker: str = "SPZ_JUMP_TABLE") -> Tuple[str, int]:
    """Returns (patched PTX, number of switches rewritten)."""
    lines = ptx.split("\n")
    n_done = 0
    pos = 0
    while True:
        at = next((i for i in range(pos, len(lines)) if marker in lines[i] and lines[i].lstrip().startswith("//")), None)
        if at is None:
            break
        m = re.search(marker + r"\s+(\d+)", lines[at])
        if not m:
            raise PatchError("marker without a case count")
        n_cases = int(m.group(1))
        labels = {mm.group(1): i for i, l in enumerate(lines) if (mm := _LABEL.match(l))}
        # the tree's root: the first setp on a 16/32-bit register against an immediate after the marker, in the same basic block
        root = None
        for i in range(at + 1, min(at + 200, len(lines))):
            if _LABEL.match(lines[i]) or _BRA.match(lines[i]) or _PBRA.match(lines[i]):
                break
            if _SETP.match(lines[i]):
                root = i
                break
        if root is None:
            raise PatchError("no compare tree after the marker")
        ms = _SETP.match(lines[root])
        sel, width = ms.group(5), ms.group(3)
        targets: List[int] = [_walk(lines, labels, root, sel, v) for v in range(n_cases)]
        if len(set(targets)) < max(2, n_cases // 2):
            raise PatchError(f"the tree after the marker reaches only {len(set(targets))} bodies for {n_cases} cases")
        # every target needs a label: reuse the one right above it, else insert one
        new_labels: Dict[int, str] = {}
        names: List[str] = []
        for t in targets:
            j = t - 1
            while j >= 0 and _SKIP.match(lines[j]):
                j -= 1
            ml = _LABEL.match(lines[j]) if j >= 0 else None
            if ml:
                names.append(ml.group(1))
            else:
                names.append(new_labels.setdefault(t, f"$L__spz_case_{n_done}_{t}"))
        tag = f"{n_done}"
        idx = f"%spz_jt_idx{tag}"
        if width == "16":
            conv = [f"\tcvt.u32.u16 \t{idx}, {sel};"]
        else:
            conv = [f"\tmov.u32 \t{idx}, {sel};"]
        repl = [f"$L__spz_jt{tag}:", "\t.branchtargets " + ", ".join(names) + ";",
                "\t{", f"\t.reg .u32 {idx};"] + conv + [f"\tbrx.idx \t{idx}, $L__spz_jt{tag};", "\t}"]
        # insert the new labels (from the bottom, so indices stay valid), then the dispatch in front of the root
        for t in sorted(new_labels, reverse=True):
            lines.insert(t, new_labels[t] + ":")
            if t <= root:
                root += 1
        lines[root:root] = repl
        pos = root + len(repl)
        n_done += 1
    return "\n".join(lines), n_done


def arms_reached(ptx: str, marker: str = "SPZ_JUMP_TABLE") -> Optional[int]:
    """Number of entries of the first rewritten table (diagnostic)."""
    m = re.search(r"\.branchtargets ([^;]*);", ptx)
    return len(m.group(1).split(",")) if m else None
