"""-m "not gpu": the batched host engine (sweep order, masks, gauge bookkeeping, jump handling, RNG replay) driven with the
numpy twin of the kernels, checked against the fixtures produced by the unmodified reference."""
import pytest

from oracle.golden_cases import CASES
from oracle.np_kernels import NumpyKernels
from tests.helpers import check_against_golden, engine_for_case

FAST = ["ising5_ref_golden", "ising10_c1_order1", "ising6_twosite_order1", "ising6_twosite_order2", "ising6_noiseless_order1",
        "ising6_noiseless_order2", "heis6_1site_order1", "ising6_1site_noiseless", "ising6_final_only_order1", "ising8_capped_order2",
        "ising8_strong_order2", "heis8_c2_order2", "ising10_c1_order2"]


@pytest.mark.parametrize("name", FAST)
def test_engine_with_twin_matches_reference(golden_dir, name):
    eng, tu, su, traj = engine_for_case(name, NumpyKernels())
    log = []
    results, diag = eng.run(tu, su, log)
    check_against_golden(name, golden_dir, results, diag, log, traj)
