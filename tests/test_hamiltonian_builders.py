"""MPO.from_pauli_sum (mpo.py:1171-1317), MPO.from_matrix (:1852-1974) and the reference's manual Hamiltonian constructor
(hamiltonian.py:49-167).  Ground truth for the operators is the dense Kronecker sum; whole runs with such Hamiltonians are pinned on
fixtures of the unmodified reference (``paulisum*`` in oracle/golden_cases.EDGE_CASES)."""
import numpy as np
import pytest

import yaqs_b200 as Y
from yaqs_b200.containers import MPO, Hamiltonian

P = {"I": np.eye(2), "X": np.array([[0, 1], [1, 0]], dtype=complex), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1.0, -1.0]).astype(complex)}


def _dense(length, terms):
    out = np.zeros((2 ** length, 2 ** length), dtype=complex)
    for coeff, labels in terms:
        m = np.ones((1, 1))
        for site in range(length):
            m = np.kron(m, P[labels.get(site, "I")])
        out += coeff * m
    return out


def test_from_pauli_sum_is_the_dense_sum_exactly():
    rng = np.random.default_rng(7)
    for _ in range(120):
        length = int(rng.integers(1, 8))
        terms, specs = [], []
        for _t in range(int(rng.integers(0, 10))):
            sites = sorted(rng.choice(length, size=int(rng.integers(0, min(length, 5) + 1)), replace=False))
            labels = {int(s): str(rng.choice(list("IXYZ"))) for s in sites}
            coeff = complex(rng.standard_normal(), rng.standard_normal())
            terms.append((coeff, labels))
            specs.append((coeff, (", " if rng.random() < 0.3 else " ").join(f"{l if rng.random() < 0.8 else l.lower()}{s}" for s, l in labels.items())))
        mpo = MPO()
        mpo.from_pauli_sum(terms=specs, length=length)
        assert mpo.length == length and [t.shape[:2] for t in mpo.tensors] == [(2, 2)] * length
        assert mpo.tensors[0].shape[2] == 1 and mpo.tensors[-1].shape[3] == 1
        assert all(a.shape[3] == b.shape[2] for a, b in zip(mpo.tensors[:-1], mpo.tensors[1:]))
        assert np.max(np.abs(mpo.to_matrix() - _dense(length, terms))) < 1e-14


def test_from_pauli_sum_shares_channels():
    """Bond dimension = idle + done + distinct remaining suffixes: 3 for the Ising chain (the textbook MPO), L - b + 1 at cut b for
    all-to-all ZZ couplings (one channel per target site, not one per pair)."""
    ising = MPO()
    ising.from_pauli_sum(terms=[(-1.0, f"Z{i} Z{i + 1}") for i in range(9)] + [(-0.5, f"X{i}") for i in range(10)], length=10)
    assert [t.shape[3] for t in ising.tensors] == [3] * 9 + [1]
    assert np.max(np.abs(ising.to_matrix() - MPO.ising(10, 1.0, 0.5).to_matrix())) == 0.0
    dense = MPO()
    dense.from_pauli_sum(terms=[(0.1 * (i + j), f"Z{i} Z{j}") for i in range(8) for j in range(i + 1, 8)], length=8)
    assert [t.shape[3] for t in dense.tensors] == [9, 8, 7, 6, 5, 4, 3, 1]
    empty = MPO()
    empty.from_pauli_sum(terms=[], length=3)
    assert np.max(np.abs(empty.to_matrix())) == 0.0


def test_from_pauli_sum_rejects_what_the_reference_rejects():
    m = MPO()
    with pytest.raises(ValueError, match="Duplicate site 1"):
        m.from_pauli_sum(terms=[(1.0, "X1 Z1")], length=3)
    with pytest.raises(ValueError, match="Invalid token"):
        m.from_pauli_sum(terms=[(1.0, "X1 Q2")], length=3)
    with pytest.raises(ValueError, match=r"Site index 3 outside \[0, 2\]"):
        m.from_pauli_sum(terms=[(1.0, "X3")], length=3)
    with pytest.raises(ValueError, match="length must be positive"):
        m.from_pauli_sum(terms=[], length=0)
    with pytest.raises(ValueError, match="physical_dimension=2"):
        m.from_pauli_sum(terms=[], length=2, physical_dimension=3)


def test_from_matrix_and_the_manual_hamiltonian_constructor():
    rng = np.random.default_rng(3)
    a = rng.standard_normal((16, 16)) + 1j * rng.standard_normal((16, 16))
    a = a + a.conj().T
    mpo = MPO.from_matrix(a)
    assert mpo.length == 4 and np.max(np.abs(mpo.to_matrix() - a)) < 1e-13
    assert np.max(np.abs(Hamiltonian(matrix=a).to_matrix() - a)) < 1e-13
    assert np.max(np.abs(Hamiltonian(4, matrix=a).to_matrix() - a)) < 1e-13
    low_rank = np.kron(P["Z"], np.kron(P["Z"], np.eye(2))) + np.kron(np.eye(2), np.kron(P["X"], P["X"]))
    assert [t.shape[3] for t in MPO.from_matrix(low_rank).tensors] == [2, 2, 1]                 # singular values <= cutoff dropped
    # tensors= takes the reference's (left, right, out, in) cores (MPO.custom transposes them, mpo.py:1146-1168)
    cores = [np.transpose(t, (2, 3, 0, 1)) for t in mpo.tensors]
    assert np.max(np.abs(Hamiltonian(tensors=cores).to_matrix() - a)) < 1e-13
    with pytest.raises(ValueError, match="exactly one of tensors, matrix, or sparse_matrix"):
        Hamiltonian(4)
    with pytest.raises(ValueError, match="exactly one of tensors, matrix, or sparse_matrix"):
        Hamiltonian(tensors=cores, matrix=a)
    with pytest.raises(ValueError, match="does not match len"):
        Hamiltonian(3, tensors=cores)
    with pytest.raises(ValueError, match="does not match physical_dimension"):
        Hamiltonian(3, matrix=a)
    import scipy.sparse
    assert np.max(np.abs(Hamiltonian(sparse_matrix=scipy.sparse.csr_matrix(low_rank)).to_matrix() - low_rank)) < 1e-13
    assert Hamiltonian.ising(4, 1.0, 0.5).length == 4 and Hamiltonian.from_mpo(mpo).mpo is mpo


def test_runs_do_not_see_the_mpo_gauge():
    """The same Ising operator as a preset, as a Pauli sum, as a dense matrix and as hand-made cores: Simulator.run gives the same
    observables (<= 1e-10) and the same bonds on the kernel twin -- H_eff depends on the operator, not on the MPO that carries it."""
    from oracle.np_kernels import NumpyKernels
    length = 5
    obs = [Y.Observable(Y.Z(), s) for s in range(length)] + [Y.Observable(Y.X(), 2)]
    noise = Y.NoiseModel([{"name": n, "sites": [i], "strength": 0.1} for i in range(length) for n in ("lowering", "pauli_z")])
    preset = Hamiltonian.ising(length, 1.0, 0.5)
    summed = MPO()
    summed.from_pauli_sum(terms=[(-1.0, f"Z{i} Z{i + 1}") for i in range(length - 1)] + [(-0.5, f"X{i}") for i in range(length)], length=length)
    variants = [preset, Hamiltonian.from_mpo(summed), Hamiltonian(matrix=preset.to_matrix()),
                Hamiltonian(tensors=[np.transpose(t, (2, 3, 0, 1)) for t in preset.mpo.tensors])]
    results = []
    for ham in variants:
        sp = Y.AnalogSimParams(observables=obs, elapsed_time=0.5, dt=0.1, num_traj=6, max_bond_dim=8, svd_threshold=1e-9, order=2, random_seed=11)
        results.append(Y.Simulator(show_progress=False, kernels=NumpyKernels()).run(Y.State(length, initial="x+"), ham, sp, noise))
    for res in results[1:]:
        for a, b in zip(results[0].expectation_values, res.expectation_values):
            assert np.max(np.abs(np.asarray(a) - np.asarray(b))) < 1e-10
        assert np.array_equal(results[0].total_bond, res.total_bond)


def test_fermi_hubbard_jordan_wigner_and_small_mpo_tools():
    """MPO.fermi_hubbard_1d(jordan_wigner=True) (mpo.py:409-470) against the dense Pauli sum of its formula (the same matrix as the
    unmodified reference's, compared in the build container: max |diff| 9e-16 at 8 orbitals); mpo_sum / +, identity, validity check."""
    t, u, length = 0.7, 1.3, 6
    terms = []
    for up in range(0, length, 2):
        terms += [(u / 4, {}), (-u / 4, {up: "Z"}), (-u / 4, {up + 1: "Z"}), (u / 4, {up: "Z", up + 1: "Z"})]
    for a in range(length - 2):
        terms += [(-t / 2, {a: p, a + 1: "Z", a + 2: p}) for p in "XY"]
    mpo = MPO.fermi_hubbard_1d(length, t, u, jordan_wigner=True)
    assert np.max(np.abs(mpo.to_matrix() - _dense(length, terms))) < 1e-14
    assert max(w.shape[3] for w in mpo.tensors) <= 7 and mpo.check_if_valid_mpo()
    assert np.max(np.abs(Hamiltonian.fermi_hubbard_1d(length, t, u, jordan_wigner=True).matrix - mpo.to_matrix())) == 0.0
    with pytest.raises(ValueError, match="even integer"):
        MPO.fermi_hubbard_1d(5, t, u, jordan_wigner=True)
    with pytest.raises(NotImplementedError):
        MPO.fermi_hubbard_1d(4, t, u)
    a, b = MPO.ising(4, 1.0, 0.5), MPO.heisenberg(4, 1.0, 0.8, 0.6, 0.1)
    total = a + b
    assert total.check_if_valid_mpo() and np.max(np.abs(total.to_matrix() - a.to_matrix() - b.to_matrix())) < 1e-14
    assert np.max(np.abs(MPO.mpo_sum([a, b, a]).to_matrix() - 2 * a.to_matrix() - b.to_matrix())) < 1e-14
    assert np.array_equal(MPO.identity(3).to_matrix(), np.eye(8))
    assert Hamiltonian.ising(3, 1.0, 0.5).sparse_matrix.shape == (8, 8)
    broken = MPO.custom([np.zeros((2, 2, 1, 3)), np.zeros((2, 2, 2, 1))])
    assert not broken.check_if_valid_mpo()
