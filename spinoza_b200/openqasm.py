"""OpenQASM 2.0 -> QuantumCircuit (mirror of /root/reference/spinoza/src/openqasm.rs).

The reference delegates lexing/parsing to the un-vendored crates `qasm ^1.0.0` and `evalexpr ^11.3.0`
(spinoza/Cargo.toml:24,26); this is a from-scratch parser for the subset the reference's importer
handles (openqasm.rs:57-165): qreg, h, x, y, z, rx, ry, rz, u, cp, cx.  Angle arguments accept
arithmetic in `pi` (the reference only does for `cp`, openqasm.rs:146; accepting it everywhere is a
superset).  Registers are laid out in declaration order (the reference iterates a HashMap,
openqasm.rs:42-53 -- SURVEY.md Q4; its fixtures are single-register).  Anything else raises, where
the reference hits `todo!()` (openqasm.rs:163-166).
"""
from __future__ import annotations

import ast
import math
import operator
import re
from pathlib import Path
from typing import Dict, List, Tuple

from .circuit import QuantumCircuit, QuantumRegister

_BIN = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
        ast.Pow: operator.pow}
_UN = {ast.UAdd: operator.pos, ast.USub: operator.neg}


def eval_angle(expr: str) -> float:
    """Evaluate an arithmetic expression over numbers and `pi` (the evalexpr context of openqasm.rs:36-40)."""

    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return float(node.value)
        if isinstance(node, ast.Name) and node.id == "pi":
            return math.pi
        if isinstance(node, ast.BinOp) and type(node.op) in _BIN:
            return _BIN[type(node.op)](ev(node.left), ev(node.right))
        if isinstance(node, ast.UnaryOp) and type(node.op) in _UN:
            return _UN[type(node.op)](ev(node.operand))
        raise ValueError(f"unsupported angle expression: {expr!r}")

    return ev(ast.parse(expr.strip().replace("^", "**"), mode="eval"))


_STMT = re.compile(r"^\s*([A-Za-z_][A-Za-z0-9_]*)\s*(?:\((.*)\))?\s*(.*)$", re.S)
_QARG = re.compile(r"^\s*([A-Za-z_][A-Za-z0-9_]*)\s*\[\s*(\d+)\s*\]\s*$")


def _strip_comments(src: str) -> str:
    return re.sub(r"//[^\n]*", "", src)


def _split_args(s: str) -> List[str]:
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def loads(qasm_as_str: str, device: int = 0, fuse: bool = True, shift: int = 0, into: QuantumCircuit = None) -> QuantumCircuit:
    """openqasm.rs:25-32.  `into`/`shift` append the program onto an existing circuit at a qubit offset."""
    stmts = [s.strip() for s in _strip_comments(qasm_as_str).split(";") if s.strip()]
    regs: Dict[str, QuantumRegister] = {}
    order: List[QuantumRegister] = []
    body: List[Tuple[str, List[str], List[str]]] = []
    for s in stmts:
        if s.startswith("OPENQASM") or s.startswith("include"):
            continue
        m = _STMT.match(s)
        if not m:
            raise ValueError(f"cannot parse statement: {s!r}")
        name, params, rest = m.group(1), m.group(2), m.group(3)
        if name == "qreg":
            q = _QARG.match(rest)
            if not q:
                raise ValueError(f"bad qreg: {s!r}")
            regs[q.group(1)] = QuantumRegister(int(q.group(2)))
            order.append(regs[q.group(1)])
            continue
        body.append((name, _split_args(params) if params is not None else [], _split_args(rest)))
    if into is None:
        qc = QuantumCircuit(*order, device=device, fuse=fuse)
    else:
        qc = into
        bits = shift
        for r in order:
            r.update_shift(bits)
            bits += len(r)

    def qubit(arg: str) -> int:
        q = _QARG.match(arg)
        if not q:
            raise ValueError(f"expected a qubit argument, got {arg!r}")
        return regs[q.group(1)][int(q.group(2))]

    for name, params, qargs in body:
        if name in ("h", "x", "y", "z"):
            for a in qargs:
                getattr(qc, name)(qubit(a))
        elif name in ("rx", "ry", "rz"):
            for a in qargs:
                getattr(qc, name)(eval_angle(params[0]), qubit(a))
        elif name == "u":
            th, ph, la = (eval_angle(p) for p in params)
            for a in qargs:
                qc.u(th, ph, la, qubit(a))
        elif name == "cp":
            qc.cp(eval_angle(params[0]), qubit(qargs[0]), qubit(qargs[1]))
        elif name == "cx":
            qc.cx(qubit(qargs[0]), qubit(qargs[1]))
        else:
            raise NotImplementedError(f"gate {name!r}: todo!() in the reference (openqasm.rs:163)")
    return qc


def load(filename, device: int = 0, fuse: bool = True) -> QuantumCircuit:
    """openqasm.rs:13-23"""
    return loads(Path(filename).read_text(), device=device, fuse=fuse)
