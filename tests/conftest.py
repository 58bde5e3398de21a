"""pytest configuration: registers the ``gpu`` marker and puts the repo root on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# keep BLAS single-threaded so CPU tests are reproducible and do not oversubscribe (reference: tests/conftest.py:21-30)
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR


# -m gpu tests that were written after round 1's GPU budget was spent and have not yet executed on a B200.
# They are collected LAST, so a first-run surprise in one of them cannot hide the 111 tests that are known green on the box
# (profiles/r01_pytest_gpu_v7_baseline_sizes.txt) behind `-x`.  Remove an entry once a profiles/ log shows it green.
FIRST_RUN_ON_GPU = ("test_backend_boundary_gpu", "test_measure_shots_replays_reference_outcomes_gpu", "ising6_scheduled", "dig_ising6_periodic_swaps",
                    "dig_ising49_c4", "test_program_matches_reference_gpu", "periodic_mpo", "test_mpo_site_apply_matches_twin", "ragged_gemm_option", "gemm_ragged_option", "cuda_dynamic_tdvp", "dig_toffoli6", "wapply_onepass_option", "dig_ising6_window_capped",
                    "test_cuda_engine_edge_cases", "test_cuda_digital_tjm_controlled_gates", "test_zz_reference_unit_kats", "heis200_c5_step")


def _gpu_unavailable() -> str | None:
    import torch
    if not torch.cuda.is_available():
        return "needs a CUDA device (run on the B200 box)"
    if not os.path.exists(os.path.join(ROOT, "yaqs_b200", "libyaqs_b200.so")):
        return "libyaqs_b200.so is not built (python -m yaqs_b200.build)"
    return None


def pytest_collection_modifyitems(config, items):
    why = _gpu_unavailable()
    if why is not None:                              # a plain `pytest tests` on a CPU box skips the GPU tests instead of failing them
        skip = pytest.mark.skip(reason=why)
        for it in items:
            if it.get_closest_marker("gpu"):
                it.add_marker(skip)
    late = [it for it in items if it.get_closest_marker("gpu") and any(tag in it.nodeid for tag in FIRST_RUN_ON_GPU)]
    if late:
        late_ids = {id(it) for it in late}
        items[:] = [it for it in items if id(it) not in late_ids] + late
