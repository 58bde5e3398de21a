"""Loader for the *unmodified* reference (munich-quantum-toolkit/yaqs) in THIS container only.

TEST INFRASTRUCTURE -- never imported by the product path (yaqs_b200/*), bench.py's GPU arm or
``-m gpu`` tests.  /root/reference does not exist on the GPU box; this module is used only to
(a) validate oracle/tjm_oracle.py against the real reference and (b) generate the committed
fixtures under tests/golden/ (see oracle/make_golden.py).

The reference cannot be imported as a package here: ``mqt/yaqs/__init__.py`` pulls in qiskit and a
hatch-generated ``_version.py`` (both absent), and ``opt_einsum`` is not installed.  We therefore
pre-register empty namespace modules so ``__init__`` never runs, and alias ``opt_einsum.contract``
to ``numpy.einsum`` (same contraction, summation order may differ at 1e-16).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("YAQS_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src", "mqt")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "yaqs"))


def load_reference() -> None:
    """Install the import shims; afterwards ``import mqt.yaqs.<submodule>`` works."""
    if "mqt.yaqs" in sys.modules and getattr(sys.modules["mqt.yaqs"], "_b200_shim", False):
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    # the reference tree is read-only by contract: no bytecode and no numba cache files next to its sources
    sys.dont_write_bytecode = True
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(os.environ.get("TMPDIR", "/tmp"), "yaqs_b200_numba_cache"))
    for name, path in (("mqt", REF_SRC), ("mqt.yaqs", os.path.join(REF_SRC, "yaqs"))):
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        mod._b200_shim = True
        sys.modules[name] = mod
    if "opt_einsum" not in sys.modules:
        oe = types.ModuleType("opt_einsum")
        oe.contract = lambda *a, **k: np.einsum(*a, optimize=True)
        sys.modules["opt_einsum"] = oe
    # digital path: qiskit is only needed for DAG handling / QASM, never for the arithmetic.
    try:
        import qiskit  # noqa: F401
    except ImportError:
        for name in ("qiskit", "qiskit.circuit", "qiskit.converters", "qiskit.dagcircuit", "qiskit.exceptions",
                     "qiskit.quantum_info", "qiskit.qasm2", "qiskit.qasm3", "qiskit.circuit.library",
                     "qiskit.quantum_info.operators", "qiskit.quantum_info.operators.predicates", "qiskit.utils", "qiskit.utils.optionals",
                     "qiskit.transpiler", "qiskit.transpiler.passes", "qiskit.providers", "qiskit.providers.fake_provider"):
            stub = types.ModuleType(name)
            stub.__path__ = []                     # a package, so that deeper stubbed submodules resolve from sys.modules
            stub.__getattr__ = lambda attr, _n=name: type(attr, (), {})  # any name resolves to a dummy class
            sys.modules[name] = stub
