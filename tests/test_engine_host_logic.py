"""-m "not gpu": the batched host engine (sweep order, masks, gauge bookkeeping, jump handling, RNG replay) driven with the
numpy twin of the kernels, checked against the fixtures produced by the unmodified reference."""
import pytest

from oracle.golden_cases import CASES
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


from oracle.golden_cases import DIGITAL_CASES  # noqa: E402
from tests.helpers import digital_engine_for_case  # noqa: E402


@pytest.mark.parametrize("name", sorted(DIGITAL_CASES))
def test_digital_engine_with_twin_matches_reference(golden_dir, name):
    """Digital TJM host logic (circuit compilation, TEBD gates, per-gate local noise, TDVP gate windows, sample barriers)
    on the numpy twin vs fixtures of the unmodified reference's digital_tjm."""
    eng, compiled, tu, traj, cfg = digital_engine_for_case(name, NumpyKernels())
    log = []
    results, diag = eng.run_digital(compiled.layers, tu, cfg["sample"], cfg["gate_mode"], log)
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
