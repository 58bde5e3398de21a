"""Observable / gate library: the matrices the device consumes as 2x2 / 4x4 constants.

Mirrors the part of the reference's ``GateLibrary`` (core/libraries/gate_library.py:129-2048) that the TJM path needs:
named one- and two-qubit gates with ``.name``, ``.matrix``, ``.interaction``, ``.sites`` / ``set_sites`` and, for the
rotation gates, the product-form generator used by TDVP gate windows.  Matrix conventions follow the reference
(qiskit little-endian remaps are handled where gates are applied, mpo_utils.py:104-160).
"""
from __future__ import annotations

import copy

import numpy as np

C = np.complex128
_I = np.eye(2, dtype=C)
_X = np.array([[0, 1], [1, 0]], dtype=C)
_Y = np.array([[0, -1j], [1j, 0]], dtype=C)
_Z = np.array([[1, 0], [0, -1]], dtype=C)


class BaseGate:
    """gate_library.py:129-567 (the fields the hot path reads)."""

    name = "custom"
    interaction = 1

    def __init__(self, matrix: np.ndarray, name: str | None = None, interaction: int | None = None) -> None:
        self.matrix = np.asarray(matrix, dtype=C)
        if name is not None:
            self.name = name
        n = int(round(np.log2(self.matrix.shape[0])))
        self.interaction = interaction if interaction is not None else n
        self.sites: list[int] | None = None
        self.generator = None

    @property
    def tensor(self) -> np.ndarray:
        if self.interaction == 1:
            return self.matrix
        return self.matrix.reshape((2,) * (2 * self.interaction))      # (out_0 .. out_{n-1}, in_0 .. in_{n-1}) in ``sites`` order

    def set_sites(self, *sites) -> None:
        flat: list[int] = []
        for s in sites:
            if isinstance(s, (list, tuple)):
                flat.extend(int(v) for v in s)
            else:
                flat.append(int(s))
        if len(flat) != self.interaction:
            raise ValueError(f"Number of sites {len(flat)} must be equal to the interaction level {self.interaction}")
        self.sites = flat

    def dag(self) -> "BaseGate":
        g = copy.deepcopy(self)
        g.matrix = self.matrix.conj().T.copy()
        return g


def _fixed(name: str, matrix, interaction: int | None = None):
    def factory() -> BaseGate:
        return BaseGate(np.array(matrix, dtype=C), name=name, interaction=interaction)
    factory.__name__ = name
    return factory


def _rot(axis: np.ndarray):
    def mat(theta: float) -> np.ndarray:
        return np.cos(theta / 2) * _I - 1j * np.sin(theta / 2) * axis
    return mat


def _rot2(pa: np.ndarray, pb: np.ndarray):
    def mat(theta: float) -> np.ndarray:
        return np.cos(theta / 2) * np.eye(4, dtype=C) - 1j * np.sin(theta / 2) * np.kron(pa, pb)
    return mat


class Entropy(BaseGate):
    """Meta-observable: bipartite entanglement entropy across a nearest-neighbour cut (gate_library.py:1873-1900)."""

    def __init__(self) -> None:
        super().__init__(np.eye(2), name="entropy", interaction=2)


class SchmidtSpectrum(BaseGate):
    """Meta-observable: the Schmidt spectrum (singular values) across a nearest-neighbour cut, a length-500 vector padded with
    NaN (gate_library.py:1903-1940, mps.py:644-678)."""

    def __init__(self) -> None:
        super().__init__(np.eye(2), name="schmidt_spectrum", interaction=2)


class PVM(BaseGate):
    """Projection-valued measurement onto a computational-basis bitstring, site 0 first (gate_library.py:1796-1870)."""

    def __init__(self, bitstring: str) -> None:
        if not isinstance(bitstring, str) or not bitstring or any(ch not in "01" for ch in bitstring):
            raise ValueError(f"PVM bitstring must be a non-empty string of 0/1, got {bitstring!r}")
        super().__init__(np.eye(2), name="pvm", interaction=1)
        self.bitstring = bitstring


class GateLibrary:
    """Namespace of gate factories; ``getattr(GateLibrary, name)`` resolves names as in the reference
    (gate_library.py:1996-2048)."""

    x = staticmethod(_fixed("x", _X))
    y = staticmethod(_fixed("y", _Y))
    z = staticmethod(_fixed("z", _Z))
    id = staticmethod(_fixed("id", _I))
    i = staticmethod(_fixed("id", _I))                 # qiskit's other names of the identity (gate_library.py:2008-2009)
    iden = staticmethod(_fixed("id", _I))
    h = staticmethod(_fixed("h", np.array([[1, 1], [1, -1]]) / np.sqrt(2)))
    sx = staticmethod(_fixed("sx", 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]])))
    sxdg = staticmethod(_fixed("sxdg", 0.5 * np.array([[1 - 1j, 1 + 1j], [1 + 1j, 1 - 1j]])))          # gate_library.py:927-946
    s = staticmethod(_fixed("s", np.array([[1, 0], [0, 1j]])))
    sdg = staticmethod(_fixed("sdg", np.array([[1, 0], [0, -1j]])))
    t = staticmethod(_fixed("t", np.array([[1, 0], [0, np.exp(1j * np.pi / 4)]])))
    tdg = staticmethod(_fixed("tdg", np.array([[1, 0], [0, np.exp(-1j * np.pi / 4)]])))
    p0 = staticmethod(_fixed("p0", np.array([[1, 0], [0, 0]])))
    p1 = staticmethod(_fixed("p1", np.array([[0, 0], [0, 1]])))
    destroy = staticmethod(_fixed("destroy", np.array([[0, 1], [0, 0]])))
    create = staticmethod(_fixed("create", np.array([[0, 0], [1, 0]])))
    xx = staticmethod(_fixed("xx", np.kron(_X, _X)))
    yy = staticmethod(_fixed("yy", np.kron(_Y, _Y)))
    zz = staticmethod(_fixed("zz", np.kron(_Z, _Z)))
    @staticmethod
    def cx() -> BaseGate:
        """gate_library.py:1148-1200: control on the first site; generator pi/4 (I - Z) (x) (I - X), which is what a long-range CX
        evolves under in the TDVP gate window (digital_tjm.py:324-452)."""
        g = BaseGate(np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]), name="cx", interaction=2)
        g.generator = [(np.pi / 4) * np.array([[0, 0], [0, 2]], dtype=C), np.array([[1, -1], [-1, 1]], dtype=C)]
        return g

    @staticmethod
    def cz() -> BaseGate:
        """gate_library.py:1203-1252: generator pi/4 (I - Z) (x) (I - Z)."""
        g = BaseGate(np.diag([1, 1, 1, -1]), name="cz", interaction=2)
        g.generator = [(np.pi / 4) * np.array([[0, 0], [0, 2]], dtype=C), np.array([[0, 0], [0, 2]], dtype=C)]
        return g

    swap = staticmethod(_fixed("swap", np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])))
    cswap = staticmethod(_fixed("cswap", np.eye(8)[[0, 1, 2, 3, 4, 6, 5, 7]], 3))         # no product-form generator (gate_library.py:1455-1499)

    @staticmethod
    def ccx() -> BaseGate:
        """Toffoli (gate_library.py:1255-1302): controls on the first two sites; generator pi/4 (I - Z) (x) P1 (x) (I - X)."""
        g = BaseGate(np.eye(8)[[0, 1, 2, 3, 4, 5, 7, 6]], name="ccx", interaction=3)
        g.generator = [(np.pi / 4) * np.array([[0, 0], [0, 2]], dtype=C), np.array([[0, 0], [0, 1]], dtype=C), np.array([[1, -1], [-1, 1]], dtype=C)]
        return g

    @staticmethod
    def ccz() -> BaseGate:
        """gate_library.py:1304-1345: generator pi/4 (I - Z) (x) P1 (x) (I - Z)."""
        g = BaseGate(np.diag([1, 1, 1, 1, 1, 1, 1, -1]), name="ccz", interaction=3)
        g.generator = [(np.pi / 4) * np.array([[0, 0], [0, 2]], dtype=C), np.array([[0, 0], [0, 1]], dtype=C), np.array([[0, 0], [0, 2]], dtype=C)]
        return g

    entropy = staticmethod(lambda: Entropy())
    schmidt_spectrum = staticmethod(lambda: SchmidtSpectrum())
    pvm = staticmethod(lambda bitstring: PVM(bitstring))

    @staticmethod
    def local(matrix) -> BaseGate:
        m = np.asarray(matrix, dtype=C)
        if m.ndim != 2 or m.shape[0] != m.shape[1]:
            raise ValueError("local observable must be a square matrix")
        return BaseGate(m, name="local", interaction=1)

    @staticmethod
    def position(*, positions) -> BaseGate:
        """One-site operator diagonal in a supplied position basis (gate_library.py:1845-1870), e.g. ``Observable("position", 3,
        positions=[-0.5, 0.5])``."""
        values = np.asarray(positions)
        if np.iscomplexobj(values):
            raise ValueError("positions must contain only real values.")
        values = np.asarray(values, dtype=np.float64)
        if values.ndim != 1 or values.size == 0:
            raise ValueError("positions must be a non-empty one-dimensional array.")
        if not np.all(np.isfinite(values)):
            raise ValueError("positions must contain only finite values.")
        g = GateLibrary.local(np.diag(values))
        g.name = "position"
        return g

    custom = BaseGate

    @staticmethod
    def rx(theta: float) -> BaseGate:
        g = BaseGate(_rot(_X)(theta), name="rx")
        g.theta = theta
        return g

    @staticmethod
    def ry(theta: float) -> BaseGate:
        g = BaseGate(_rot(_Y)(theta), name="ry")
        g.theta = theta
        return g

    @staticmethod
    def rz(theta: float) -> BaseGate:
        g = BaseGate(_rot(_Z)(theta), name="rz")
        g.theta = theta
        return g

    @staticmethod
    def p(theta: float) -> BaseGate:
        g = BaseGate(np.array([[1, 0], [0, np.exp(1j * theta)]]), name="p")
        g.theta = theta
        return g

    @staticmethod
    def u(theta: float, phi: float, lam: float) -> BaseGate:
        m = np.array([[np.cos(theta / 2), -np.exp(1j * lam) * np.sin(theta / 2)],
                      [np.exp(1j * phi) * np.sin(theta / 2), np.exp(1j * (phi + lam)) * np.cos(theta / 2)]])
        return BaseGate(m, name="u")

    @staticmethod
    def u3(theta: float, phi: float, lam: float) -> BaseGate:
        """qiskit's older name of ``u`` (gate_library.py:2015)."""
        return GateLibrary.u(theta, phi, lam)

    @staticmethod
    def u2(phi: float, lam: float) -> BaseGate:
        """gate_library.py:1074-1107: (1 / sqrt 2) [[1, -e^{i lam}], [e^{i phi}, e^{i (phi + lam)}]]."""
        g = BaseGate(np.array([[1, -np.exp(1j * lam)], [np.exp(1j * phi), np.exp(1j * (phi + lam))]]) / np.sqrt(2), name="u2")
        g.phi, g.lam = phi, lam
        return g

    @staticmethod
    def u1(theta: float) -> BaseGate:
        """The phase gate under its older name (gate_library.py:2017: ``u1 = Phase``)."""
        return GateLibrary.p(theta)

    @staticmethod
    def _two_rot(name: str, pa, pb, theta: float) -> BaseGate:
        g = BaseGate(_rot2(pa, pb)(theta), name=name, interaction=2)
        g.theta = theta
        # exp(-i theta/2 A (x) B): generator factors with theta/2 folded into the first one (gate_library.py:1546,1608,1670)
        g.generator = [(theta / 2) * pa, pb.copy()]
        return g

    @staticmethod
    def rxx(theta: float) -> BaseGate:
        return GateLibrary._two_rot("rxx", _X, _X, theta)

    @staticmethod
    def ryy(theta: float) -> BaseGate:
        return GateLibrary._two_rot("ryy", _Y, _Y, theta)

    @staticmethod
    def rzz(theta: float) -> BaseGate:
        return GateLibrary._two_rot("rzz", _Z, _Z, theta)

    @staticmethod
    def cp(theta: float) -> BaseGate:
        g = BaseGate(np.diag([1, 1, 1, np.exp(1j * theta)]), name="cp", interaction=2)
        g.theta = theta
        g.generator = [-theta * np.array([[0, 0], [0, 1]], dtype=C), np.array([[0, 0], [0, 1]], dtype=C)]      # gate_library.py:1397-1401
        return g


# Class-style aliases used as ``Observable(Z(), site)`` in the reference's examples and tests.
X, Y, Z, Id, H = GateLibrary.x, GateLibrary.y, GateLibrary.z, GateLibrary.id, GateLibrary.h
XX, YY, ZZ = GateLibrary.xx, GateLibrary.yy, GateLibrary.zz
P0, P1, Destroy, Create = GateLibrary.p0, GateLibrary.p1, GateLibrary.destroy, GateLibrary.create
Rx, Ry, Rz, Rxx, Ryy, Rzz = GateLibrary.rx, GateLibrary.ry, GateLibrary.rz, GateLibrary.rxx, GateLibrary.ryy, GateLibrary.rzz
