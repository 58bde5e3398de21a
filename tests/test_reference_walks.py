"""Short differential walks with the UNMODIFIED reference on the other side -- only where its source tree is present (the build container;
the GPU box has no /root/reference, and this test is not ``gpu``-marked).  The walks run in a process of their own because the import
shim that loads the reference registers stand-in modules (qiskit, opt_einsum) that must not leak into the test process.  The long walks
(hundreds of draws each) are run by hand with the same entry points (``python -m oracle.run_reference_walks``)."""
import os
import subprocess
import sys

import pytest

from oracle.ref_shim import reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not reference_available(), reason="needs the reference source tree (build container only)")
def test_walks_against_the_unmodified_reference():
    """Seven walks, ten draws each: analog / digital / gate-list / program draws reference against oracle, and the three public calls
    (analog Simulator.run, program Simulator.run, circuits) reference against yaqs_b200 on the kernel twin."""
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    proc = subprocess.run([sys.executable, "-m", "oracle.run_reference_walks", "--cases", "10", "--seed", "5"], cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=900)
    lines = [ln for ln in proc.stdout.splitlines() if "clamped" not in ln]
    report = "\n".join(ln for ln in lines if ln.startswith("==") or "mismatches" in ln or "BAD" in ln)
    assert proc.returncode == 0 and "ALL WALKS CLEAN" in proc.stdout, f"{report}\n{proc.stderr[-1500:]}"
    assert sum("0 mismatches" in ln for ln in lines) == 7, report
