"""OpenQASM 2 front end without qiskit.

The reference hands OpenQASM text or a path to ``qiskit.qasm2`` (digital/utils/qasm_utils.py:133-162).  qiskit is not part of the B200
boxes' image, so this module reads the OpenQASM 2 subset the reference can simulate at all -- the gate names of
``SUPPORTED_QISKIT_GATE_NAMES`` (digital/utils/dag_utils.py:51-83, plus ``cu1`` = ``cp`` and ``u0`` dropped as in qelib1.inc), ``barrier``,
``measure`` (the compiler drops terminal measurements and rejects others, digital_tjm.py:268-286), several ``qreg`` declarations laid out in
declaration order, whole-register broadcast -- into a ``yaqs_b200.Circuit``.  ``gate`` / ``opaque`` definitions, ``if``, ``reset`` and
OpenQASM 3 are refused with a message naming the statement; the entry point and its error messages mirror ``load_circuit``.
"""
from __future__ import annotations

import ast
import math
import operator
import os
import re
from pathlib import Path

from .circuits import _FIXED_GATES, _PARAM_GATES, Circuit

_VERSION = re.compile(r"OPENQASM\s+(\d+)", re.IGNORECASE)
_INVALID_INPUT = ("Expected a QuantumCircuit, a filesystem path to an OpenQASM file, or a raw OpenQASM string whose first substantive line "
                  "declares OPENQASM.")
_ALIASES = {"cu1": "cp", "CX": "cx", "U": "u"}
_FUNCS = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "exp": math.exp, "ln": math.log, "sqrt": math.sqrt}
_BINARY = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv, ast.Pow: operator.pow}


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.DOTALL)
    return "\n".join(line.split("//", maxsplit=1)[0] for line in text.splitlines())


def qasm_version(text: str) -> str | None:
    """"2" / "3" when the first substantive line declares OPENQASM, else None (qasm_utils.py:85-100)."""
    for line in _strip_comments(text).splitlines():
        line = line.strip()
        if not line:
            continue
        if line.upper().startswith("OPENQASM"):
            m = _VERSION.match(line)
            return "3" if (m is not None and m.group(1) == "3") else "2"
        return None
    return None


def _angle(expr: str) -> float:
    """A parameter expression of OpenQASM 2: numbers, pi, + - * / ^, unary minus, parentheses, sin cos tan exp ln sqrt."""
    def walk(node) -> float:
        if isinstance(node, ast.Expression):
            return walk(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
            return float(node.value)
        if isinstance(node, ast.Name) and node.id == "pi":
            return math.pi
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            return -walk(node.operand) if isinstance(node.op, ast.USub) else walk(node.operand)
        if isinstance(node, ast.BinOp) and type(node.op) in _BINARY:
            return _BINARY[type(node.op)](walk(node.left), walk(node.right))
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _FUNCS and len(node.args) == 1 and not node.keywords:
            return _FUNCS[node.func.id](walk(node.args[0]))
        raise ValueError(f"unsupported parameter expression '{expr.strip()}'")
    try:
        return walk(ast.parse(expr.strip().replace("^", "**"), mode="eval"))
    except SyntaxError as exc:
        raise ValueError(f"unsupported parameter expression '{expr.strip()}'") from exc


def _split_params(text: str) -> list[str]:
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
            continue
        depth += ch == "("
        depth -= ch == ")"
        cur += ch
    return parts + [cur] if cur.strip() else parts


def parse_openqasm2(text: str) -> Circuit:
    regs: dict[str, tuple[int, int]] = {}                  # name -> (first wire, size)
    ops: list[tuple] = []
    width = 0

    def wires(arg: str) -> list[int]:
        m = re.fullmatch(r"([A-Za-z_]\w*)\s*(?:\[\s*(\d+)\s*\])?", arg.strip())
        if m is None or m.group(1) not in regs:
            raise ValueError(f"unknown quantum argument '{arg.strip()}'")
        first, size = regs[m.group(1)]
        if m.group(2) is None:
            return list(range(first, first + size))
        if int(m.group(2)) >= size:
            raise ValueError(f"index out of range in '{arg.strip()}'")
        return [first + int(m.group(2))]

    for raw in _strip_comments(text).split(";"):
        stmt = " ".join(raw.split())
        if not stmt:
            continue
        head = stmt.split("(")[0].split(" ")[0]
        if head.upper() == "OPENQASM" or head == "include" or head == "creg":
            continue
        if head == "qreg":
            m = re.fullmatch(r"qreg ([A-Za-z_]\w*)\s*\[\s*(\d+)\s*\]", stmt)
            if m is None:
                raise ValueError(f"malformed statement '{stmt}'")
            regs[m.group(1)] = (width, int(m.group(2)))
            width += int(m.group(2))
            continue
        if head in ("gate", "opaque", "if", "reset") or stmt.startswith("if("):
            raise ValueError(f"OpenQASM statement '{head}' is not supported by the qiskit-free reader (the TJM compiler rejects reset and "
                             "classically controlled operations, and simulates library gates only)")
        if head == "measure":
            for q in wires(stmt[len("measure"):].split("->")[0]):
                ops.append(("measure", (q,), ()))
            continue
        if head == "barrier":
            qs = [q for a in stmt[len("barrier"):].split(",") for q in wires(a)]
            ops.append(("barrier", tuple(qs), ()))
            continue
        m = re.fullmatch(r"([A-Za-z_]\w*)\s*(?:\((.*)\))?\s*(.*)", stmt)
        if m is None:
            raise ValueError(f"malformed statement '{stmt}'")
        name = _ALIASES.get(m.group(1), m.group(1))
        params = tuple(_angle(p) for p in _split_params(m.group(2))) if m.group(2) else ()
        if name == "u0":
            continue                                       # an idle cycle in qelib1.inc
        if name not in _FIXED_GATES and name not in _PARAM_GATES:
            raise ValueError(f"unsupported gate '{name}'")
        if len(params) != _PARAM_GATES.get(name, 0):
            raise ValueError(f"gate '{name}' takes {_PARAM_GATES.get(name, 0)} parameter(s), got {len(params)}")
        args = [wires(a) for a in m.group(3).split(",")]
        span = max(len(a) for a in args)
        if any(len(a) not in (1, span) for a in args):
            raise ValueError(f"register sizes do not match in '{stmt}'")
        for k in range(span):                              # whole-register arguments broadcast
            qubits = tuple(a[k] if len(a) > 1 else a[0] for a in args)
            if len(set(qubits)) != len(qubits):
                raise ValueError(f"duplicate qubit arguments in '{stmt}'")
            ops.append((name, qubits, params))
    if width == 0:
        raise ValueError("the OpenQASM program declares no qreg")
    circ = Circuit(width)
    for name, qubits, params in ops:
        circ._add(name, qubits, params)
    return circ


def load_circuit(circuit):
    """``load_circuit`` of the reference (qasm_utils.py:133-162): a circuit object unchanged, a ``Path`` / path string to an OpenQASM
    file, or raw OpenQASM source whose first substantive line declares OPENQASM."""
    if not isinstance(circuit, (str, os.PathLike)):
        return circuit
    if isinstance(circuit, str) and qasm_version(circuit) is not None:
        text = circuit
    else:
        path = Path(circuit)
        if not path.is_file():
            raise ValueError(_INVALID_INPUT)
        text = path.read_text(encoding="utf-8")
        if qasm_version(text) is None:
            raise ValueError(_INVALID_INPUT)
    if qasm_version(text) == "3":
        raise ImportError("OpenQASM 3 loading requires qiskit with qiskit-qasm3-import; the qiskit-free reader of yaqs_b200 takes OpenQASM 2.")
    return parse_openqasm2(text)
