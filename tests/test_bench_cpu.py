"""bench.py pieces that run without a GPU: the CPU reference arm end to end (tiny chain) and the SURVEY 8(d) flop model."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--length", "8", "--chi", "4", "--steps", "1",
                          "--warmup", "1", "--cpu-cores", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "trajectory-steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "C3" and d["vs_baseline"] is None


def test_flop_model_matches_survey_headline():
    """Uniform chi = 128, D = 3, 5.08 matvecs per Krylov call, no jumps: SURVEY.md 8(d) quotes F_step ~ 2.15e12 (2 x 1.0e12 for the
    sweeps + 396 gauge shifts of 3.4e8)."""
    from bench import flop_model_step
    L = 100
    bonds = [1.0] + [128.0] * (L - 1) + [1.0]
    mpo = [1.0] + [3.0] * (L - 1) + [1.0]
    fm = flop_model_step(bonds, mpo, 2 * (2 * L - 3) * 5.08, 2 * (2 * L - 4) * 5.08, 0.0)
    assert 1.9e12 < fm["total"] < 2.4e12
    assert abs(sum(v for k, v in fm.items() if k != "total") - fm["total"]) < 1e-6 * fm["total"]
    assert 1.1e12 < fm["krylov"] < 1.4e12 and 1.0e11 < fm["gauge_shifts"] < 1.6e11


def test_cpu_extrapolation_is_linear_in_chain_length():
    """The reference arm samples two short saturated chains and extrapolates the half-step time linearly in the chain length
    (bench.cpu_extrapolate); chains that are not longer than the second sample are timed as they are."""
    from bench import CPU_SAMPLE_LENGTHS, cpu_extrapolate
    a, b = CPU_SAMPLE_LENGTHS
    v, t = cpu_extrapolate({"length": 100}, {a: 4.0, b: 9.0}, 16)
    assert abs(t - (9.0 + (100 - b) * 5.0 / (b - a))) < 1e-12 and abs(v - 0.5 * 16 / t) < 1e-12
    v, t = cpu_extrapolate({"length": 10}, {10: 0.5}, 4)
    assert t == 0.5 and v == 4.0


def test_hbm_byte_model():
    """15 S + 4 E bytes per trajectory-step; C1 (L = 10, chi <= 16, D = 3) is quoted as ~1.7 MB with all bonds at 16."""
    from bench import hbm_bytes_step
    bonds = [1.0] + [16.0] * 9 + [1.0]
    mpo = [1.0] + [3.0] * 9 + [1.0]
    by = hbm_bytes_step(bonds, mpo)
    assert 1.2e6 < by < 2.0e6


def test_reference_arm_of_the_digital_workload():
    """--workload C4 (BASELINE.json configs[3]): the CPU arm runs the oracle's digital_tjm over the first W + K Trotter steps on every worker."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C4", "--length", "5", "--steps", "1",
                          "--warmup", "1", "--cpu-cores", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["config"]["workload"] == "C4" and d["value"] > 0 and d["cpu_baseline"]["cores"] == 2
    assert "Trotter" in d["metric"] and d["e2e"]["value"] == d["value"]
