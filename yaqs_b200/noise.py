"""Host-side noise model: normalised process dictionaries that the engine flattens into device tables.

Mirrors the reference's ``NoiseModel`` surface (core/data_structures/noise_model.py:227-665) and operator library
(core/libraries/noise_library.py:22-196): same process keys (``name``, ``sites``, ``strength``, ``matrix`` | ``factors``),
same validation errors, same ``sample`` semantics for distribution-valued strengths, same ``is_pauli`` rule.
Scheduled jumps are a SURVEY.md 8(f) "next" row and are rejected at run time.
"""
from __future__ import annotations

import copy
import logging
import math
import re
from typing import Any

import numpy as np

logger = logging.getLogger(__name__)

PAULI = {
    "x": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    "z": np.array([[1, 0], [0, -1]], dtype=np.complex128),
}
_LOWERING = np.array([[0, 1], [0, 0]], dtype=np.complex128)      # noise_library.py:22-37
_RAISING = np.array([[0, 0], [1, 0]], dtype=np.complex128)       # noise_library.py:38-53

OPERATORS: dict[str, np.ndarray] = {
    "lowering": _LOWERING, "raising": _RAISING,
    "pauli_x": PAULI["x"], "pauli_y": PAULI["y"], "pauli_z": PAULI["z"],
    "x": PAULI["x"], "y": PAULI["y"], "z": PAULI["z"],
    "lowering_two": np.kron(_LOWERING, _LOWERING), "raising_two": np.kron(_RAISING, _RAISING),
}
for _a in "xyz":
    for _b in "xyz":
        OPERATORS[f"crosstalk_{_a}{_b}"] = np.kron(PAULI[_a], PAULI[_b])

_CROSSTALK = re.compile(r"^(longrange_)?crosstalk_([xyz]{2})$")
_DISTRIBUTIONS = {"normal", "lognormal", "truncated_normal"}


def get_operator(name: str) -> np.ndarray:
    """noise_model.py:561-586."""
    m = _CROSSTALK.fullmatch(name)
    if m:
        return np.kron(PAULI[m.group(2)[0]], PAULI[m.group(2)[1]]).copy()
    if name not in OPERATORS:
        raise ValueError(f"Unknown noise operator '{name}'. Supported fixed names: {', '.join(sorted(OPERATORS))}. "
                         "Also accepted: names matching crosstalk_[xyz]{2} or longrange_crosstalk_[xyz]{2}.")
    return OPERATORS[name].copy()


def _square(value: object, label: str) -> np.ndarray:
    try:
        arr = np.array(value, dtype=np.complex128, copy=True)
    except (TypeError, ValueError) as exc:
        raise TypeError(f"{label} must be a numeric array.") from exc
    if arr.ndim != 2 or arr.shape[0] != arr.shape[1]:
        raise ValueError(f"{label} must be a square 2-D array, got shape {arr.shape}.")
    if not np.all(np.isfinite(arr)):
        raise ValueError(f"{label} entries must be finite.")
    return arr


def _real(value: object, label: str, nonneg: bool) -> float:
    if isinstance(value, bool) or not isinstance(value, (int, float, np.floating, np.integer)):
        raise TypeError(f"{label} must be a real number (booleans are not allowed).")
    x = float(value)
    if not math.isfinite(x):
        raise ValueError(f"{label} must be finite, got {x}.")
    if nonneg and x < 0:
        raise ValueError(f"{label} must be nonnegative (got {x}). Standard TJM/MCWF jump probabilities require nonnegative rates; "
                         "YAQS does not currently implement signed density-matrix generators or non-Markovian reverse-jump unravelings.")
    return x


def _strength(value: object):
    if isinstance(value, dict):
        unknown = set(value) - {"distribution", "mean", "std"}
        if unknown:
            raise ValueError(f"Unknown distribution keys: {sorted(unknown)}. Supported keys: ['distribution', 'mean', 'std'].")
        if "distribution" not in value:
            raise ValueError("Noise strength dict must contain 'distribution' key.")
        if value["distribution"] not in _DISTRIBUTIONS:
            raise ValueError(f"Unsupported distribution type: {value['distribution']}. Supported: {sorted(_DISTRIBUTIONS)}.")
        mean = _real(value.get("mean", 0.0), "distribution mean", False)
        std = _real(value.get("std", 0.0), "distribution std", False)
        if std < 0:
            raise ValueError(f"distribution std must be nonnegative, got {std}.")
        return {"distribution": value["distribution"], "mean": mean, "std": std}
    return _real(value, "process strength", True)


def _sites(sites: object, kind: str) -> list[int]:
    if not isinstance(sites, (list, tuple)):
        raise TypeError(f"{kind} 'sites' must be a list or tuple of integers.")
    if len(sites) not in (1, 2):
        raise ValueError(f"{kind} must have exactly 1 or 2 sites, got {len(sites)}.")
    out = []
    for s in sites:
        if isinstance(s, bool) or not isinstance(s, (int, np.integer)):
            raise TypeError(f"{kind} site indices must be integers (booleans are not allowed).")
        if int(s) < 0:
            raise ValueError(f"{kind} site indices must be nonnegative, got {int(s)}.")
        out.append(int(s))
    if len(out) == 2 and out[0] == out[1]:
        raise ValueError(f"{kind} two-site indices must be distinct, got {out}.")
    return out


def _unit_phase_match(mat: np.ndarray, ref: np.ndarray) -> bool:
    """noise_model.py:617-630."""
    if mat.shape != ref.shape:
        return False
    idx = np.unravel_index(int(np.argmax(np.abs(ref))), ref.shape)
    if abs(ref[idx]) < 1e-14 or abs(mat[idx]) < 1e-14:
        return bool(np.allclose(mat, ref, atol=1e-10, rtol=0.0))
    phase = mat[idx] / ref[idx]
    if not np.isclose(abs(phase), 1.0, atol=1e-10, rtol=0.0):
        return False
    return bool(np.allclose(mat, phase * ref, atol=1e-10, rtol=0.0))


_P1 = tuple(PAULI[k] for k in "xyz")
_P2 = tuple(np.kron(a, b) for a in _P1 for b in _P1)


def is_pauli(proc: dict[str, Any]) -> bool:
    """True iff L is a Pauli (string) up to a unit phase, i.e. L^dagger L = 1 and the dissipator is a scalar
    (noise_model.py:644-665)."""
    sites = proc["sites"]
    if len(sites) == 1:
        return "matrix" in proc and any(_unit_phase_match(np.asarray(proc["matrix"], dtype=np.complex128), q) for q in _P1)
    if len(sites) != 2:
        return False
    if abs(sites[1] - sites[0]) == 1 and "matrix" in proc:
        return any(_unit_phase_match(np.asarray(proc["matrix"], dtype=np.complex128), q) for q in _P2)
    if abs(sites[1] - sites[0]) > 1 and "factors" in proc:
        return all(any(_unit_phase_match(np.asarray(f, dtype=np.complex128), q) for q in _P1) for f in proc["factors"])
    return False


class NoiseModel:
    """Collection of Lindblad jump processes (noise_model.py:227-299)."""

    def __init__(self, processes: list[dict[str, Any]] | None = None, scheduled_jumps: list[dict[str, Any]] | None = None) -> None:
        self.processes: list[dict[str, Any]] = []
        self.scheduled_jumps: list[dict[str, Any]] = []
        if scheduled_jumps is not None:
            if not isinstance(scheduled_jumps, (list, tuple)):
                raise TypeError("scheduled_jumps must be a list or tuple of dictionaries.")
            self.scheduled_jumps = [self._normalize_scheduled_jump(j) for j in scheduled_jumps]
        if processes is None:
            return
        if not isinstance(processes, (list, tuple)):
            raise TypeError("processes must be a list or tuple of dictionaries.")
        self.processes = [self._normalize(p) for p in processes]

    @staticmethod
    def get_operator(name: str) -> np.ndarray:
        return get_operator(name)

    @staticmethod
    def _normalize_scheduled_jump(jump: object) -> dict[str, Any]:
        """noise_model.py:298-339: {time, sites, name[, matrix]} -> sites ascending, 'matrix' resolved."""
        if not isinstance(jump, dict):
            raise TypeError("Each scheduled jump must be a dictionary.")
        for key in ("time", "sites", "name"):
            if key not in jump:
                raise ValueError(f"Each scheduled jump must have a '{key}' key.")
        out = dict(jump)
        if "factors" in out:
            raise ValueError("Scheduled jumps do not accept 'factors'; use 'matrix' for custom operators.")
        if not isinstance(out["name"], str):
            raise TypeError("Scheduled jump 'name' must be a string.")
        if not out["name"]:
            raise ValueError("Scheduled jump 'name' must be a nonempty string.")
        out["time"] = _real(out["time"], "Scheduled jump time", nonneg=False)
        sites = _sites(out["sites"], "Scheduled jump")
        user_matrix = "matrix" in out
        swapped = False
        if len(sites) == 2:
            ordered = sorted(sites)
            swapped = ordered != sites
            if abs(ordered[1] - ordered[0]) != 1:
                raise ValueError(f"Scheduled jump acts on non-adjacent sites {sites}. Only nearest-neighbor scheduled jumps are supported.")
            if swapped and user_matrix:
                raise ValueError(f"Custom full scheduled-jump matrices require ascending site order; got sites {sites}.")
            sites = ordered
        out["sites"] = sites
        if user_matrix:
            out["matrix"] = _square(out["matrix"], "Scheduled jump matrix")
        else:
            cross = _CROSSTALK.fullmatch(out["name"])
            if cross:
                a, b = cross.group(2)
                if swapped:
                    a, b = b, a
                out["matrix"] = np.kron(PAULI[a], PAULI[b]).copy()
            else:
                out["matrix"] = get_operator(out["name"])
        return out

    @staticmethod
    def _normalize(original: object) -> dict[str, Any]:
        """noise_model.py:452-490 (+ :316-450)."""
        if not isinstance(original, dict):
            raise TypeError("Each noise process must be a dictionary.")
        for key in ("name", "sites", "strength"):
            if key not in original:
                raise ValueError(f"Each process must have a '{key}' key.")
        proc = dict(original)
        name = proc["name"]
        if not isinstance(name, str):
            raise TypeError("Process 'name' must be a string.")
        if not name:
            raise ValueError("Process 'name' must be a nonempty string.")
        proc["strength"] = _strength(proc["strength"])
        sites = _sites(proc["sites"], "Process")
        has_matrix, has_factors = "matrix" in original, "factors" in original
        if has_factors and original["factors"] is None:
            raise ValueError("Process 'factors' must be a sequence of exactly two square matrices, not None.")
        if has_matrix and has_factors:
            raise ValueError("Process cannot specify both 'matrix' and 'factors'.")
        if len(sites) == 1:
            if has_factors:
                raise ValueError("One-site processes do not accept 'factors'.")
            proc["sites"] = sites
            proc["matrix"] = _square(proc["matrix"], "Process matrix") if has_matrix else get_operator(name)
            return proc
        ordered = sorted(sites)
        swapped = ordered != sites
        if swapped and has_matrix:
            raise ValueError(f"Custom full two-site matrices require ascending site order; got sites {sites}. "
                             "Use ascending sites or supply 'factors'.")
        proc["sites"] = ordered
        cross = _CROSSTALK.fullmatch(name)
        letters = None
        if cross:
            a, b = cross.group(2)
            letters = (b, a) if swapped else (a, b)
        if ordered[1] - ordered[0] != 1:
            if has_matrix:
                raise ValueError("Non-adjacent two-site processes require 'factors' (a full 'matrix' embedding is not accepted here).")
            if not has_factors:
                if letters is None:
                    raise ValueError("Non-adjacent 2-site processes must specify 'factors' unless named "
                                     "crosstalk_[xyz]{2} or longrange_crosstalk_[xyz]{2}.")
                proc["factors"] = (PAULI[letters[0]].copy(), PAULI[letters[1]].copy())
            else:
                f = original["factors"]
                if not isinstance(f, (list, tuple)) or len(f) != 2:
                    raise ValueError("Process 'factors' must be a sequence of exactly two square matrices.")
                left, right = _square(f[0], "Process factor[0]"), _square(f[1], "Process factor[1]")
                proc["factors"] = (right, left) if swapped else (left, right)
            proc.pop("matrix", None)
            return proc
        if has_factors:
            raise ValueError("Adjacent two-site processes use 'matrix', not 'factors'.")
        if has_matrix:
            proc["matrix"] = _square(proc["matrix"], "Process matrix")
        elif letters is not None:
            proc["matrix"] = np.kron(PAULI[letters[0]], PAULI[letters[1]]).copy()
        else:
            proc["matrix"] = get_operator(name)
        proc.pop("factors", None)
        return proc

    def sample(self, rng: np.random.Generator | int | None = None) -> "NoiseModel":
        """Resolve distribution-valued strengths into one realisation of static disorder (noise_model.py:492-559)."""
        gen = np.random.default_rng(rng)
        out = []
        for proc in self.processes:
            new = copy.deepcopy(proc)
            s = proc["strength"]
            if isinstance(s, dict):
                kind, mean, std = s["distribution"], s["mean"], s["std"]
                if kind == "normal":
                    val = float(gen.normal(loc=mean, scale=std))
                    if val < 0:
                        logger.warning("Sampled noise strength %f using 'normal' distribution (mean=%f, std=%f) was negative and clamped to 0.0.",
                                       val, mean, std)
                    val = float(max(0.0, val))
                elif kind == "lognormal":
                    val = float(gen.lognormal(mean=mean, sigma=std))
                else:
                    if math.isclose(std, 0.0, abs_tol=1e-8):
                        val = float(max(0.0, mean))
                    else:
                        from scipy.stats import truncnorm
                        val = float(truncnorm.rvs((0.0 - mean) / std, np.inf, loc=mean, scale=std, random_state=gen))
                new["strength"] = _real(val, "sampled process strength", True)
            else:
                new["strength"] = _real(s, "process strength", True)
            out.append(new)
        model = object.__new__(NoiseModel)
        model.processes = out
        model.scheduled_jumps = copy.deepcopy(self.scheduled_jumps)
        return model
