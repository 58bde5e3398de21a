"""Qiskit-free circuit description + the reference's layer compilation for the digital TJM.

The reference takes a qiskit ``QuantumCircuit`` and schedules it with ``circuit_to_dag`` / ``dag.front_layer()``
(digital/digital_tjm.py:142-184, 233-301).  qiskit is not available on the B200 boxes, so this module provides
  * ``Circuit``: a minimal gate list with the qiskit-style builder methods the YAQS circuit library uses
    (``rx, ry, rz, h, x, ..., rzz, rxx, ryy, cx, cz, swap, cp, barrier(label=...)``),
  * ``compile_circuit``: the same front-layer scheduling restated on that list -- per layer: single-qubit gates, then
    multi-qubit gates split into "even" / "odd" groups by their lowest qubit index, labelled ``SAMPLE_OBSERVABLES``
    barriers as sample points, other barriers dropped, terminal measurements dropped, and
  * ``create_ising_circuit`` (core/libraries/circuit_library.py:28-79).
A real ``qiskit.QuantumCircuit`` is accepted too when qiskit is importable (converted through ``circuit.data``).
"""
from __future__ import annotations

from dataclasses import dataclass, field

from .gates import BaseGate, GateLibrary

SAMPLE_LABEL = "SAMPLE_OBSERVABLES"
_PARAM_GATES = {"rx": 1, "ry": 1, "rz": 1, "p": 1, "u1": 1, "rxx": 1, "ryy": 1, "rzz": 1, "cp": 1, "u2": 2, "u": 3, "u3": 3}      # name -> number of angles
_FIXED_GATES = {"x", "y", "z", "h", "id", "i", "iden", "sx", "sxdg", "s", "sdg", "t", "tdg", "cx", "cz", "swap", "ccx", "ccz", "cswap"}


@dataclass
class Op:
    name: str
    qubits: tuple[int, ...]
    params: tuple[float, ...] = ()
    label: str | None = None


@dataclass
class Circuit:
    """Ordered gate list on ``num_qubits`` wires (the subset of qiskit's builder API that the YAQS libraries use)."""
    num_qubits: int
    ops: list[Op] = field(default_factory=list)

    def _add(self, name, qubits, params=(), label=None):
        for q in qubits:
            if not 0 <= q < self.num_qubits:
                raise ValueError(f"qubit index {q} out of range for a {self.num_qubits}-qubit circuit")
        self.ops.append(Op(name, tuple(int(q) for q in qubits), tuple(float(p) for p in params), label))
        return self

    def barrier(self, *qubits, label: str | None = None):
        return self._add("barrier", qubits if qubits else tuple(range(self.num_qubits)), label=label)

    def measure_all(self):
        for q in range(self.num_qubits):
            self._add("measure", (q,))
        return self

    def __getattr__(self, name):
        if name in _FIXED_GATES:
            return lambda *qubits: self._add(name, qubits)
        if name in _PARAM_GATES:
            nparam = _PARAM_GATES[name]

            def add(*args, **kw):
                vals = list(args) + [kw[k] for k in ("theta", "phi", "lam") if k in kw]
                qs = [kw[k] for k in ("qubit", "qubit1", "qubit2") if k in kw]
                params, qubits = vals[:nparam], vals[nparam:] + qs
                return self._add(name, qubits, params)
            return add
        raise AttributeError(name)


@dataclass(frozen=True)
class CompiledLayer:
    """digital_tjm.py:49-60."""
    single_qubit_gates: tuple
    even_two_qubit_gates: tuple
    odd_two_qubit_gates: tuple
    sample_points: int


@dataclass(frozen=True)
class CompiledCircuit:
    layers: tuple
    num_mid_measurements: int


def _to_gate(op: Op) -> BaseGate:
    """convert_dag_to_tensor_algorithm (digital/utils/dag_utils.py:305): gate matrix + sites."""
    factory = getattr(GateLibrary, op.name, None)
    if factory is None:
        raise ValueError(f"unsupported gate '{op.name}'")
    gate = factory(*op.params) if op.params else factory()
    gate.set_sites(*op.qubits)
    return gate


def from_qiskit(qc) -> Circuit:
    circ = Circuit(qc.num_qubits)
    for inst in qc.data:
        op = inst.operation
        qubits = tuple(qc.find_bit(q).index for q in inst.qubits)
        circ._add(op.name, qubits, tuple(float(p) for p in getattr(op, "params", ())), getattr(op, "label", None))
    return circ


def compile_circuit(circuit) -> CompiledCircuit:
    """_compile_circuit + process_layer (digital_tjm.py:142-184, 233-301) on a gate list."""
    if not isinstance(circuit, Circuit):
        circuit = from_qiskit(circuit)
    ops = list(circuit.ops)
    alive = [True] * len(ops)
    wires: list[list[int]] = [[] for _ in range(circuit.num_qubits)]      # op indices per wire, program order
    for k, op in enumerate(ops):
        for q in op.qubits:
            wires[q].append(k)
    head = [0] * circuit.num_qubits

    def front() -> list[int]:
        out = []
        for q in range(circuit.num_qubits):
            while head[q] < len(wires[q]) and not alive[wires[q][head[q]]]:
                head[q] += 1
        seen = set()
        for q in range(circuit.num_qubits):
            if head[q] < len(wires[q]):
                k = wires[q][head[q]]
                if k not in seen and all(wires[w][head[w]] == k for w in ops[k].qubits):
                    seen.add(k)
                    out.append(k)
        return sorted(out)

    def is_terminal_measure(k: int) -> bool:
        return not any(alive[j] and set(ops[j].qubits) & set(ops[k].qubits) for j in range(k + 1, len(ops)))

    layers, mid = [], 0
    while any(alive):
        single, even, odd, samples = [], [], [], []
        for k in front():
            op = ops[k]
            if op.name == "measure":
                if not is_terminal_measure(k):
                    raise ValueError("Non-terminal measure operations are not supported during simulation; "
                                     "removing them would ignore state collapse and classical dependencies.")
                alive[k] = False
            elif op.name == "barrier":
                if op.label is not None and str(op.label).upper() == SAMPLE_LABEL:
                    samples.append(k)
                else:
                    alive[k] = False
            elif len(op.qubits) == 1:
                single.append(k)
            elif min(op.qubits) % 2 == 0:
                even.append(k)
            else:
                odd.append(k)
        single.sort(key=lambda k: ops[k].qubits[0])
        even.sort(key=lambda k: min(ops[k].qubits))
        odd.sort(key=lambda k: min(ops[k].qubits))
        for k in (*single, *even, *odd, *samples):
            alive[k] = False
        mid += len(samples)
        if single or even or odd or samples:
            layers.append(CompiledLayer(tuple(_to_gate(ops[k]) for k in single), tuple(_to_gate(ops[k]) for k in even),
                                        tuple(_to_gate(ops[k]) for k in odd), len(samples)))
    return CompiledCircuit(tuple(layers), mid)


def create_ising_circuit(L: int, J: float, g: float, dt: float, timesteps: int, *, periodic: bool = False,
                         sample_every_step: bool = False) -> Circuit:
    """core/libraries/circuit_library.py:28-79 (same gate order, a full barrier after every gate).  ``sample_every_step``
    adds a labelled SAMPLE_OBSERVABLES barrier after every Trotter step but the last."""
    alpha, beta = -2 * dt * g, -2 * dt * J
    circ = Circuit(L)
    for step in range(timesteps):
        for site in range(L):
            circ.rx(alpha, site)
            circ.barrier()
        for site in range(L // 2):
            circ.rzz(beta, 2 * site, 2 * site + 1)
            circ.barrier()
        for site in range(1, L // 2):
            circ.rzz(beta, 2 * site - 1, 2 * site)
            circ.barrier()
        if L % 2 != 0 and L != 1:
            circ.rzz(beta, L - 2, L - 1)
            circ.barrier()
        if periodic and L > 1:
            circ.rzz(beta, 0, L - 1)
            circ.barrier()
        if sample_every_step and step < timesteps - 1:
            circ.barrier(label=SAMPLE_LABEL)
    return circ


# ----------------------------------------------------------------------------------------------------------------------
# The other Trotter-circuit builders of the reference's library (core/libraries/circuit_library.py:82-349), qiskit-free.
# Each is written as (bond list per interaction layer) x (one emitting loop); tests/test_circuit_library.py holds the gate lists the
# unmodified reference's builders emit and compares them gate by gate, barriers included.
# ----------------------------------------------------------------------------------------------------------------------
def _chain_bonds(length: int) -> list[tuple[int, int]]:
    """Even pairs (0,1), (2,3), ..., then odd pairs (1,2), (3,4), ..., then the tail pair an odd chain leaves over."""
    bonds = [(2 * k, 2 * k + 1) for k in range(length // 2)] + [(2 * k - 1, 2 * k) for k in range(1, length // 2)]
    if length % 2 != 0 and length != 1:
        bonds.append((length - 2, length - 1))
    return bonds


def _snake(num_rows: int, num_cols: int):
    """Row-major snake through the grid (even rows left to right, odd rows right to left): neighbours in a row stay neighbours
    on the chain, vertical neighbours are at most 2 * num_cols - 1 apart."""
    return lambda row, col: row * num_cols + (col if row % 2 == 0 else num_cols - 1 - col)


def _grid_bond_groups(num_rows: int, num_cols: int):
    """Bond lines of the 2-D Heisenberg builder, group by group: (horizontal: even bonds of every row, then odd bonds of every row;
    vertical: even bonds of every column, then odd bonds of every column).  (The 2-D Ising builder walks line by line instead.)"""
    at = _snake(num_rows, num_cols)
    horizontal = [[(at(r, c), at(r, c + 1)) for c in range(start, num_cols - 1, 2)] for start in (0, 1) for r in range(num_rows)]
    vertical = [[(at(r, c), at(r + 1, c)) for r in range(start, num_rows - 1, 2)] for start in (0, 1) for c in range(num_cols)]
    return horizontal, vertical


def create_heisenberg_circuit(L: int, Jx: float, Jy: float, Jz: float, h: float, dt: float, timesteps: int, *,   # noqa: N803
                              periodic: bool = False) -> Circuit:
    """circuit_library.py:153-226: per Trotter step rz(-2 dt h) on every site, then the ZZ, XX and YY layers over even pairs, odd
    pairs, the odd-chain tail pair and (periodic) the pair (0, L-1).  The reference places three barriers, kept where it has them:
    behind the periodic ZZ gate, and behind the XX tail pair of an odd chain."""
    circ = Circuit(L)
    wrap = periodic and L > 1
    odd_tail = L % 2 != 0 and L != 1
    for _ in range(timesteps):
        for site in range(L):
            circ.rz(-2 * dt * h, site)
        for name, coupling in (("rzz", Jz), ("rxx", Jx), ("ryy", Jy)):
            for a, b in _chain_bonds(L):
                getattr(circ, name)(-2 * dt * coupling, a, b)
            if name == "rxx" and odd_tail:
                circ.barrier()
            if wrap:
                getattr(circ, name)(-2 * dt * coupling, 0, L - 1)
                if name == "rzz":
                    circ.barrier()
    return circ


def create_2d_ising_circuit(num_rows: int, num_cols: int, J: float, g: float, dt: float, timesteps: int) -> Circuit:   # noqa: N803
    """circuit_library.py:82-150: the grid mapped onto the chain along a snake; per Trotter step rx(-2 dt g) on every qubit, then
    rzz(-2 dt J) row by row (even bonds of the row, then its odd bonds) and column by column, a full barrier behind every rzz."""
    circ = Circuit(num_rows * num_cols)
    at = _snake(num_rows, num_cols)
    for _ in range(timesteps):
        for row in range(num_rows):
            for col in range(num_cols):
                circ.rx(-2 * dt * g, at(row, col))
        lines = [[(at(r, c), at(r, c + 1)) for start in (0, 1) for c in range(start, num_cols - 1, 2)] for r in range(num_rows)]
        lines += [[(at(r, c), at(r + 1, c)) for start in (0, 1) for r in range(start, num_rows - 1, 2)] for c in range(num_cols)]
        for line in lines:
            for a, b in line:
                circ.rzz(-2 * dt * J, a, b)
                circ.barrier()
    return circ


def create_2d_heisenberg_circuit(num_rows: int, num_cols: int, Jx: float, Jy: float, Jz: float, h: float, dt: float,   # noqa: N803
                                 timesteps: int) -> Circuit:
    """circuit_library.py:229-349: snake ordering; per Trotter step rz(-2 dt h) everywhere, then for ZZ, XX, YY in turn the four bond
    groups horizontal-even, horizontal-odd, vertical-even, vertical-odd (no barriers)."""
    circ = Circuit(num_rows * num_cols)
    at = _snake(num_rows, num_cols)
    horizontal, vertical = _grid_bond_groups(num_rows, num_cols)
    for _ in range(timesteps):
        for row in range(num_rows):
            for col in range(num_cols):
                circ.rz(-2.0 * dt * h, at(row, col))
        for name, coupling in (("rzz", Jz), ("rxx", Jx), ("ryy", Jy)):
            for line in horizontal + vertical:
                for a, b in line:
                    getattr(circ, name)(-2.0 * dt * coupling, a, b)
    return circ


def create_1d_fermi_hubbard_circuit(L: int, u: float, t: float, mu: float, num_trotter_steps: int, dt: float, timesteps: int) -> Circuit:   # noqa: N803
    """circuit_library.py:352-415: H = -mu/2 (I - Z) + u/4 (I - Z)(I - Z) - t/2 (XX + YY) on two registers (spin-up = wires 0 .. L-1,
    spin-down = wires L .. 2L-1).  One second-order Trotter slice is  chemical potential - on-site interaction - hopping - on-site
    interaction - chemical potential,  repeated ``num_trotter_steps * timesteps`` times: p(mu dt / 2n) on every wire, cp(-u dt / 2n)
    between the two spins of a site, and rxx / ryy(-t dt / n) on the even bonds, then the odd bonds, of each register (the hop is
    written from the higher site to the lower one)."""
    n = num_trotter_steps
    circ = Circuit(2 * L)
    up, down = range(L), range(L, 2 * L)

    def chemical_potential():
        for j in range(L):
            circ.p(mu * dt / (2 * n), up[j])
            circ.p(mu * dt / (2 * n), down[j])

    def onsite():
        for j in range(L):
            circ.cp(-u * dt / (2 * n), up[j], down[j])

    def hopping():
        for parity in (0, 1):
            for j in range(parity, L - 1, 2):
                for reg in (up, down):
                    circ.rxx(-dt * t / n, reg[j + 1], reg[j])
                    circ.ryy(-dt * t / n, reg[j + 1], reg[j])

    for _ in range(n * timesteps):
        chemical_potential()
        onsite()
        hopping()
        onsite()
        chemical_potential()
    return circ


def _u3_angles(unitary) -> tuple[float, float, float]:
    """(theta, phi, lambda) of u(theta, phi, lambda) equal to a 2 x 2 unitary up to a global phase (circuit_library_utils.py:32-64):
    the phase is fixed by making the (0, 0) entry real and non-negative."""
    import numpy as np
    m = np.asarray(unitary, dtype=np.complex128) * np.exp(-1j * np.angle(unitary[0, 0]))
    theta = 2 * np.arccos(np.clip(np.abs(m[0, 0]), -1.0, 1.0))
    if np.isclose(float(np.sin(theta / 2)), 0.0):
        return float(theta), 0.0, float(np.angle(m[1, 1]) - np.angle(m[0, 0]))
    return float(theta), float(np.angle(m[1, 0])), float(np.angle(-m[0, 1]))


def nearest_neighbour_random_circuit(n_qubits: int, layers: int, seed: int = 42) -> Circuit:
    """circuit_library.py:637-680 (the prescription of arXiv:2002.07730): per layer a random rotation exp(-i theta n.sigma) on every qubit
    -- theta and the azimuth uniform in [0, 2 pi), the polar angle uniform in [0, pi), drawn in that order from ``default_rng(seed)`` and
    written as a ``u`` gate -- then cz or cx (a fair coin per pair) on the pairs (1,2), (3,4), ... in even layers and (0,1), (2,3), ... in
    odd layers, and a barrier."""
    import numpy as np
    rng = np.random.default_rng(seed)
    circ = Circuit(n_qubits)
    for layer in range(layers):
        for qubit in range(n_qubits):
            theta, alpha, phi = rng.uniform(0, 2 * np.pi), rng.uniform(0, np.pi), rng.uniform(0, 2 * np.pi)
            axis = np.array([np.sin(alpha) * np.cos(phi), np.sin(alpha) * np.sin(phi), np.cos(alpha)])
            generator = (axis[0] * np.array([[0, 1], [1, 0]]) + axis[1] * np.array([[0, -1j], [1j, 0]]) + axis[2] * np.array([[1, 0], [0, -1]]))
            circ.u(*_u3_angles(np.cos(theta) * np.eye(2) - 1j * np.sin(theta) * generator), qubit)     # (n.sigma)^2 = 1
        for a in range(1 if layer % 2 == 0 else 0, n_qubits - 1, 2):
            if rng.random() < 0.5:
                circ.cz(a, a + 1)
            else:
                circ.cx(a, a + 1)
        circ.barrier()
    return circ


def _pauli_string_hop(circ: Circuit, i: int, j: int, alpha: float) -> None:
    """exp(-i alpha/2 (X_i Z...Z X_j)) exp(-i alpha/2 (Y_i Z...Z Y_j)) in elementary gates (circuit_library.py:442-517): for each of the
    two strings, rotate the end qubits into the Z basis (ry(pi/2) for X, rx(pi/2) for Y), fold the parity of qubits i .. j-1 onto qubit
    j with a CX ladder (highest control first), rz(alpha) on qubit j, and undo ladder and basis change in reverse."""
    import numpy as np
    if i >= j:
        raise IndexError("Assumption i < j violated.")
    for basis in ("ry", "rx"):
        rotate = getattr(circ, basis)
        rotate(np.pi / 2, i)
        rotate(np.pi / 2, j)
        for k in reversed(range(i, j)):
            circ.cx(k, j)
        circ.rz(alpha, j)
        for k in range(i, j):
            circ.cx(k, j)
        rotate(-np.pi / 2, i)
        rotate(-np.pi / 2, j)


def create_2d_fermi_hubbard_circuit(Lx: int, Ly: int, u: float, t: float, mu: float, num_trotter_steps: int, dt: float,   # noqa: N803
                                    timesteps: int) -> Circuit:
    """circuit_library.py:520-634: lattice site p = y Lx + x holds qubits 2p (spin up) and 2p + 1 (spin down); one second-order slice is
    chemical potential p(-mu dt / 2n) - on-site cp(-u dt / 2n) - hopping - on-site - chemical potential.  The hopping term walks the
    horizontal bonds starting at even x, then odd x, then the vertical bonds starting at even y, then odd y, and hops both spins over
    each bond as Jordan-Wigner strings (``_pauli_string_hop`` with alpha = t dt / n)."""
    n = num_trotter_steps
    sites = Lx * Ly
    circ = Circuit(2 * sites)

    def chemical_potential():
        for q in range(2 * sites):
            circ.p(-mu * dt / (2 * n), q)

    def onsite():
        for p in range(sites):
            circ.cp(-u * dt / (2 * n), 2 * p, 2 * p + 1)

    bonds = [(y * Lx + x, y * Lx + x + 1) for parity in (0, 1) for y in range(Ly) for x in range(parity, Lx - 1, 2)]
    bonds += [(y * Lx + x, (y + 1) * Lx + x) for parity in (0, 1) for y in range(parity, Ly - 1, 2) for x in range(Lx)]
    for _ in range(timesteps * n):
        chemical_potential()
        onsite()
        for p1, p2 in bonds:
            for spin in (0, 1):
                _pauli_string_hop(circ, 2 * p1 + spin, 2 * p2 + spin, t * dt / n)
        onsite()
        chemical_potential()
    return circ
