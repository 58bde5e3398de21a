"""-m "not gpu": the batched host engine (sweep order, masks, gauge bookkeeping, jump handling, RNG replay) driven with the
numpy twin of the kernels, checked against the fixtures produced by the unmodified reference."""
import pytest

from oracle.golden_cases import CASES, EDGE_CASES
from oracle.np_kernels import NumpyKernels
from tests.helpers import check_against_golden, engine_for_case

FAST = ["ising5_ref_golden", "ising10_c1_order1", "ising6_twosite_order1", "ising6_twosite_order2", "ising6_noiseless_order1",
        "ising6_noiseless_order2", "heis6_1site_order1", "ising6_1site_noiseless", "ising6_final_only_order1", "ising8_capped_order2",
        "ising8_strong_order2", "heis8_c2_order2", "ising10_c1_order2", "ising12_chi40_gram", "ising6_entropy_order2", "heis6_entropy_order1",
        "ising6_pvm_order2", "ising6_scheduled_order1", "ising6_scheduled_only_order1"]


@pytest.mark.parametrize("name", FAST)
def test_engine_with_twin_matches_reference(golden_dir, name):
    eng, tu, su, traj = engine_for_case(name, NumpyKernels())
    log = []
    results, diag = eng.run(tu, su, log)
    check_against_golden(name, golden_dir, results, diag, log, traj)


@pytest.mark.parametrize("name", sorted(EDGE_CASES))
@pytest.mark.parametrize("gram_min_k", [33, 0])
def test_engine_with_twin_edge_cases(golden_dir, name, gram_min_k):
    """Edge cases of the domain against the unmodified reference: a two-site and a three-site chain, max_bond_dim = 1, the truncation
    modes "relative", "hard_cutoff", "relative_discarded_weight" inside whole runs, a threshold that truncates at almost every split --
    on the Jacobi route and with the Gram / Cholesky-QR routes forced onto these sizes."""
    eng, tu, su, traj = engine_for_case(name, NumpyKernels(), gram_min_k=gram_min_k)
    log = []
    results, diag = eng.run(tu, su, log)
    assert int(eng.status.abs().sum()) == 0
    check_against_golden(name, golden_dir, results, diag, log, traj)


@pytest.mark.parametrize("name", ["ising10_c1_order2", "heis8_c2_order2", "ising8_strong_order2", "ising10_c1_order1", "heis6_1site_order1"])
def test_engine_with_twin_gram_and_cholqr_routes(golden_dir, name):
    """The Gram / eigen-solver SVD route and the Cholesky-QR route of the keep-all gauge shifts (csrc/yb_cholqr.cuh), forced onto the
    small fixtures (production takes them from 2 chi >= 33): same jumps, identical bonds, observables <= 1e-10; the shifts of
    rank-deficient bonds (product states, bonds right after a lowering jump) must come back through the fallback."""
    eng, tu, su, traj = engine_for_case(name, NumpyKernels(), gram_min_k=0)
    log = []
    results, diag = eng.run(tu, su, log)
    assert eng.stats["cholqr_calls"] > 0 and eng.stats["gram_svd_calls"] > 0
    assert 0 < eng.stats["cholqr_fallbacks"] < eng.stats["cholqr_calls"] * eng.B
    check_against_golden(name, golden_dir, results, diag, log, traj)


@pytest.mark.parametrize("name", ["heis8_c2_order2", "ising10_c1_order2"])
def test_engine_with_twin_capped_split_keeps_top_rows_only(golden_dir, name):
    """Capped splits on the Gram route back-transform only the eigenvector rows they can keep (yb_heev_rows_top / ROTATE_TOP /
    yb_svd_finalize_top): same jumps, identical bonds, observables <= 1e-10 against the unmodified reference."""
    eng, tu, su, traj = engine_for_case(name, NumpyKernels(), gram_min_k=0)
    log = []
    results, diag = eng.run(tu, su, log)
    assert eng.stats["heev_top_calls"] > 0
    check_against_golden(name, golden_dir, results, diag, log, traj)


from oracle.golden_cases import CONTROLLED_DIGITAL_CASES, DIGITAL_CASES  # noqa: E402
from tests.helpers import digital_engine_for_case  # noqa: E402


@pytest.mark.parametrize("name", sorted(DIGITAL_CASES))
def test_digital_engine_with_twin_matches_reference(golden_dir, name):
    """Digital TJM host logic (circuit compilation, TEBD gates, per-gate local noise, TDVP gate windows, sample barriers)
    on the numpy twin vs fixtures of the unmodified reference's digital_tjm."""
    eng, compiled, tu, traj, cfg = digital_engine_for_case(name, NumpyKernels())
    log = []
    results, diag = eng.run_digital(compiled.layers, tu, cfg["sample"], cfg["gate_mode"], log)
    check_against_golden(name, golden_dir, results, diag, log, traj)


@pytest.mark.parametrize("name", sorted(CONTROLLED_DIGITAL_CASES))
def test_digital_engine_with_twin_controlled_gates(golden_dir, name):
    """cx / cz / cp with control and target in both orders, adjacent and across spectators, through all four gate routes -- in the
    TDVP modes they evolve under the product-form generators the reference gives them (gate_library.py:1193-1197, 1244-1248,
    1397-1401) -- and the single-qubit names u2 / u3 / u1 / sxdg / i of the qiskit name table; fixtures of the unmodified reference."""
    eng, compiled, tu, traj, cfg = digital_engine_for_case(name, NumpyKernels())
    log = []
    results, diag = eng.run_digital(compiled.layers, tu, cfg["sample"], cfg["gate_mode"], log)
    assert int(eng.status.abs().sum()) == 0
    check_against_golden(name, golden_dir, results, diag, log, traj)


def test_compile_circuit_matches_reference_layering():
    """One gate per layer behind full barriers; even / odd groups by lowest qubit; labelled barriers become sample points;
    without barriers the front layer packs commuting gates (digital_tjm.py:142-184, 233-301)."""
    from yaqs_b200.circuits import Circuit, compile_circuit, create_ising_circuit
    cc = compile_circuit(create_ising_circuit(5, 1.0, 0.5, 0.1, 2, sample_every_step=True))
    assert cc.num_mid_measurements == 1
    kinds = [(len(l.single_qubit_gates), len(l.even_two_qubit_gates), len(l.odd_two_qubit_gates), l.sample_points) for l in cc.layers]
    per_step = [(1, 0, 0, 0)] * 5 + [(0, 1, 0, 0)] * 2 + [(0, 0, 1, 0)] + [(0, 0, 1, 0)]
    assert kinds == per_step + [(0, 0, 0, 1)] + per_step
    c = Circuit(4)
    c.h(0); c.h(1); c.cx(0, 1); c.cx(2, 3); c.rzz(0.3, 1, 2); c.measure_all()
    cc = compile_circuit(c)
    kinds = [(len(l.single_qubit_gates), len(l.even_two_qubit_gates), len(l.odd_two_qubit_gates)) for l in cc.layers]
    assert kinds == [(2, 1, 0), (0, 1, 0), (0, 0, 1)]
    assert cc.layers[0].even_two_qubit_gates[0].sites == [2, 3]


def test_dynamic_tdvp_engine_matches_reference_below_the_cap(golden_dir):
    """tdvp_mode="dynamic" (integrators.py:294-511) with no cap: every site takes the 2-site branch with the uncapped split; same
    jumps, identical bonds, observables <= 1e-10 against the unmodified reference."""
    name = "heis6_dynamic_nocap_order2"
    eng, tu, su, traj = engine_for_case(name, NumpyKernels())
    log = []
    results, diag = eng.run(tu, su, log)
    check_against_golden(name, golden_dir, results, diag, log, traj)


DYNAMIC_EXACT = [(6, 8, 8), (6, 4, 8), (8, 16, 16), (6, 2, 2)]


@pytest.mark.parametrize("length,chi0,cap", DYNAMIC_EXACT)
def test_dynamic_tdvp_engine_is_exact_at_full_bond_dimension(length, chi0, cap):
    check_dynamic_exact(NumpyKernels(), length, chi0, cap)


def check_dynamic_exact(kernels, length, chi0, cap):
    """Where bonds sit at the cap the reference's right-to-left hand-over evolves the Householder factor R where R^T belongs
    (decompositions.py:82 + integrators.py:458), which is neither gauge invariant nor a TDVP step: at FULL bond dimension, where
    TDVP has no projection error, the reference (restated literally by the oracle, pinned to its fixtures) misses exp(-i dt H)|psi>
    while its own 2-site mode is exact.  The engine takes the projector-splitting step instead and must be exact there; per-trajectory
    branch masks are exercised by batching states whose bonds reach the cap at different sites."""
    import numpy as np
    import scipy.linalg as sl

    from oracle import tjm_oracle as O
    from yaqs_b200.engine import BatchedTJM, EngineParams
    mpo = O.ising_mpo(length, 1.0, 0.8)
    ham = O.mpo_to_matrix(mpo)
    states = [O.random_mps(length, c, 3 + k) for k, c in enumerate((chi0, max(chi0 // 2, 1), chi0))]
    ep = EngineParams(dt=0.1, n_steps=1, order=1, max_bond_dim=cap, svd_threshold=1e-14, krylov_tol=1e-12, tdvp_mode="dynamic")
    eng = BatchedTJM(kernels, length, len(states), mpo, ep, None, [], [s.tensors for s in states])
    eng.tdvp(eng.phi)
    assert int(eng.status.max().item()) == 0
    for b, s in enumerate(states):
        exact = sl.expm(-1j * 0.1 * ham) @ O.mps_to_vector(s)
        got = O.mps_to_vector(O.OracleMPS(eng.phi.download(b), 0))
        if cap == 2:                      # the end bonds themselves sit at the cap: the sweep structure skips the outer half steps there
            assert abs(np.vdot(got, got) - 1.0) < 1e-10
            continue
        assert abs(abs(np.vdot(got, exact)) - 1.0) < 1e-9, (b, abs(np.vdot(got, exact)))
    if cap != 2:                          # the reference's own numbers, for the record: far from exact once a bond is at the cap
        p = O.OracleParams(dt=0.1, n_steps=1, order=1, max_bond_dim=cap, svd_threshold=1e-14, krylov_tol=1e-12, tdvp_mode="dynamic")
        st = states[0].copy()
        O.tdvp(st, mpo, p)
        ref_overlap = abs(np.vdot(O.mps_to_vector(st), sl.expm(-1j * 0.1 * ham) @ O.mps_to_vector(states[0])))
        assert ref_overlap < 0.9999 if chi0 == cap else True


@pytest.mark.parametrize("mode", ["2site", "dynamic"])
def test_tdvp_sweeps_greater_than_one_matches_oracle(mode):
    """tdvp_sweeps = 3 (tdvp.py:24-66: three symmetric substeps of dt / 3): engine on the numpy twin against the oracle restatement
    (pinned to the reference) on an uncapped chain, both sweep geometries that take a sweep plan."""
    import numpy as np

    from oracle import tjm_oracle as O
    from yaqs_b200.engine import BatchedTJM, EngineParams
    length = 6
    init = O.random_mps(length, 2, 11)
    mpo = O.heisenberg_mpo(length, 1.0, 0.7, 0.4, 0.2)
    p = O.OracleParams(dt=0.2, n_steps=1, order=1, max_bond_dim=None, svd_threshold=1e-10, krylov_tol=1e-8, tdvp_mode=mode, tdvp_sweeps=3)
    st = init.copy()
    O.tdvp(st, mpo, p)
    ep = EngineParams(dt=0.2, n_steps=1, order=1, max_bond_dim=None, svd_threshold=1e-10, krylov_tol=1e-8, tdvp_mode=mode, tdvp_sweeps=3)
    eng = BatchedTJM(NumpyKernels(), length, 2, mpo, ep, None, [], init.tensors)
    eng.tdvp(eng.phi)
    got = eng.phi.download(1)
    assert [t.shape[2] for t in got] == [t.shape[2] for t in st.tensors]
    want = O.mps_to_vector(st)
    vec = O.mps_to_vector(O.OracleMPS(got, 0))
    assert abs(np.vdot(want, vec) - np.vdot(want, want)) < 1e-10 * 10 and abs(np.vdot(vec, vec) - np.vdot(want, want)) < 1e-9


def test_in_kernel_philox_equals_philox_tables():
    """rng_mode="philox": drawing the uniforms inside yb_jump_test (BatchedTJM.philox) gives bit for bit the run that reads them from the
    tables yb_philox_uniform fills -- trajectory stream positions and the per-time-step sample streams included."""
    import numpy as np
    from yaqs_b200.rng import philox_source, philox_tables
    name = "ising10_c1_order2"
    K = NumpyKernels()
    eng_t, _, _, traj = engine_for_case(name, K)
    tu, su = philox_tables(K, 1234, traj, eng_t.p.n_times, eng_t.p.order, K.device)
    log_t = []
    res_t, diag_t = eng_t.run(tu, su, log_t)
    eng_k, _, _, _ = engine_for_case(name, K)
    eng_k.philox = philox_source(1234, traj, K.device)
    log_k = []
    res_k, diag_k = eng_k.run(None, None, log_k)
    assert eng_t.stats["jumps"] > 0
    assert np.array_equal(res_t.numpy(), res_k.numpy()) and np.array_equal(diag_t.numpy(), diag_k.numpy())
    assert len(log_t) == len(log_k)
    for a, b in zip(log_t, log_k):
        assert np.array_equal(a["u1"], b["u1"]) and np.array_equal(a["jumped"], b["jumped"])
