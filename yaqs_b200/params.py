"""Simulation parameters and observables: the configuration surface of the reference, kept name for name.

``Observable``              core/data_structures/simulation_parameters.py:360-416
observable ordering         :419-456  (workers evaluate in site-sorted order; Result is in user order)
``SIMULATION_PRESETS``      :46-51
``AnalogSimParams``         :477-613
``DigitalSimParams``        :616-750
"""
from __future__ import annotations

import copy

import numpy as np

from .gates import BaseGate, GateLibrary

SIMULATION_PRESETS = {
    "fast": {"svd_threshold": 1e-3, "max_bond_dim": 16, "num_traj": 128, "krylov_tol": 1e-3},
    "balanced": {"svd_threshold": 1e-6, "max_bond_dim": 128, "num_traj": 256, "krylov_tol": 1e-4},
    "accurate": {"svd_threshold": 1e-9, "max_bond_dim": 4096, "num_traj": 1024, "krylov_tol": 1e-6},
    "exact": {"svd_threshold": 1e-13, "max_bond_dim": None, "num_traj": 1024, "krylov_tol": 1e-12},
}
_USE_PRESET = object()
_TRUNC_MODES = ("discarded_weight", "relative", "hard_cutoff", "relative_discarded_weight")


class Observable:
    """What to measure: a gate (or named gate, or raw one-site matrix) and the site(s)."""

    def __init__(self, gate, sites=None, **gate_kwargs) -> None:
        if isinstance(gate, str):
            if gate != "pvm" and hasattr(GateLibrary, gate):
                resolved = getattr(GateLibrary, gate)(**gate_kwargs)
            else:                                     # simulation_parameters.py:388-400: any other string is a PVM bitstring
                if gate_kwargs:
                    raise TypeError(f"Unknown observable {gate!r} does not accept observable parameters.")
                resolved = GateLibrary.pvm(gate)
        elif isinstance(gate, BaseGate):
            if gate_kwargs:
                raise TypeError("Observable parameters are only supported for named observables.")
            resolved = gate
        else:
            if gate_kwargs:
                raise TypeError("Observable parameters are only supported for named observables.")
            resolved = GateLibrary.local(gate)
        self.gate = copy.deepcopy(resolved)
        if resolved.name != "pvm":                    # a PVM acts on the whole chain and carries no sites (:413-416)
            assert sites is not None
            self.sites = sites
            self.gate.set_sites(self.sites)


def prepare_observable_ordering(observables):
    """Stable sort by first site (:419-456): returns (sorted list, user index -> sorted row)."""
    if not observables:
        return [], ()
    site_obs = [p for p in enumerate(observables) if p[1].gate.name != "pvm"]
    pvm_obs = [p for p in enumerate(observables) if p[1].gate.name == "pvm"]          # appended in their original order
    pairs = sorted(site_obs, key=lambda p: ((p[1].sites[0] if isinstance(p[1].sites, list) else p[1].sites), p[0])) + pvm_obs
    user_to_sorted = [0] * len(observables)
    for row, (user_i, _) in enumerate(pairs):
        user_to_sorted[user_i] = row
    return [o for _, o in pairs], tuple(user_to_sorted)


def _check_observable_mix(observables) -> None:
    """simulation_parameters.py:583,725."""
    assert all(o.gate.name == "pvm" for o in observables) or all(o.gate.name != "pvm" for o in observables), (
        "We currently have not implemented mixed observable and projective-measurement simulation.")


def engine_observables(sorted_observables, length: int):
    """The sorted observables as the batched engine takes them: (matrix, sites) for local one-/two-site operators,
    ("entropy" | "schmidt_spectrum", [i, i+1]) and ("pvm:<bitstring>", []) for the meta-observables (mps.py:1200-1217)."""
    out = []
    for o in sorted_observables:
        name = o.gate.name
        if name == "pvm":
            assert len(o.gate.bitstring) == length, "Bitstring length must match number of sites"
            out.append(("pvm:" + o.gate.bitstring, []))
        elif name in ("entropy", "schmidt_spectrum"):
            sites = o.sites
            assert isinstance(sites, list), "Given metric requires a list of sites"
            assert len(sites) == 2, "Given metric requires 2 sites to act on."
            assert max(sites) - min(sites) == 1, "Entropy and Schmidt cuts must be nearest neighbor."
            for s in sites:
                assert s in range(length), f"Observable acting on non-existing site: {s}"
            out.append((name, [min(sites), max(sites)]))
        else:
            if o.gate.interaction not in (1, 2):
                raise ValueError("Local observable must be one-site or nearest-neighbor two-site.")
            where = [o.sites] if isinstance(o.sites, (int, np.integer)) else list(o.sites)
            for s in where:                               # (the reference trips over such a site inside its centre shifts, also with an AssertionError)
                assert 0 <= s < length, f"Observable acting on non-existing site: {s}"
            out.append((np.asarray(o.gate.matrix, dtype=np.complex128), where))
    return out


class _Ordering:
    observables: list

    @property
    def sorted_observables(self):
        return prepare_observable_ordering(self.observables)[0]

    @property
    def observable_sorted_indices(self):
        return prepare_observable_ordering(self.observables)[1]


def _validate_seed(seed) -> None:
    if seed is None:
        return
    if isinstance(seed, bool) or not isinstance(seed, int):
        raise TypeError(f"random_seed must be int or None, got {type(seed).__name__}.")
    if seed < 0:
        raise ValueError(f"random_seed must be non-negative, got {seed}.")


def _time_grid(elapsed_time, dt) -> int:
    """:92-160: elapsed_time must be an integer multiple of dt."""
    for name, v in (("elapsed_time", elapsed_time), ("dt", dt)):
        if isinstance(v, bool) or not isinstance(v, (int, float, np.floating, np.integer)):
            raise TypeError(f"{name} must be a real number, got {type(v).__name__}.")
    e, d = float(elapsed_time), float(dt)
    if not np.isfinite(e):
        raise ValueError(f"elapsed_time must be finite, got {elapsed_time!r}.")
    if not np.isfinite(d):
        raise ValueError(f"dt must be finite, got {dt!r}.")
    if d <= 0.0:
        raise ValueError(f"dt must be positive, got {d}.")
    if e < 0.0:
        raise ValueError(f"elapsed_time must be non-negative, got {e}.")
    if not e > 0.0:
        return 0
    n = round(e / d)
    residual = abs(e - n * d)
    tol = min(max(np.spacing(e), abs(n) * np.spacing(d), 8 * np.finfo(np.float64).eps * max(e, n * d, d)), 0.25 * d)
    if n <= 0 or residual > tol:
        raise ValueError(f"elapsed_time ({e}) must be an integer multiple of dt ({d}); got elapsed_time/dt = {e / d} "
                         f"(nearest integer {n}, time residual {residual}).")
    return n


def _preset_values(preset) -> dict:
    if preset not in SIMULATION_PRESETS:
        raise ValueError(f"preset must be one of {sorted(SIMULATION_PRESETS)!r}, got {preset!r}.")
    return SIMULATION_PRESETS[preset]


def _resolve_max_bond_dim(max_bond_dim, preset_value):
    """:308-327: an int, None, or left to the preset.  (Like the reference this stores any int; a cap below one is refused when a
    run builds its EngineParams, where the reference would fail inside the first split.)"""
    if max_bond_dim is _USE_PRESET:
        return preset_value
    if max_bond_dim is None or isinstance(max_bond_dim, int):
        return max_bond_dim
    raise TypeError(f"max_bond_dim must be int, None, or omitted, got {type(max_bond_dim).__name__}.")


def _numerics(self, pv: dict, num_traj, max_bond_dim, trunc_mode, svd_threshold, krylov_tol) -> None:
    """The block both constructors share, checked in the reference's order (:597-604, :744-750)."""
    self.num_traj = num_traj if num_traj is not None else pv["num_traj"]
    self.max_bond_dim = _resolve_max_bond_dim(max_bond_dim, pv["max_bond_dim"])
    if not isinstance(trunc_mode, str) or trunc_mode not in _TRUNC_MODES:
        raise ValueError(f"trunc_mode must be one of {sorted(_TRUNC_MODES)!r}, got {trunc_mode!r}.")
    self.trunc_mode = trunc_mode
    thr = float(svd_threshold if svd_threshold is not None else pv["svd_threshold"])
    if not np.isfinite(thr) or thr < 0.0:
        raise ValueError(f"svd_threshold must be a finite non-negative float, got {thr!r}.")
    self.svd_threshold = thr
    kt = float(krylov_tol if krylov_tol is not None else pv["krylov_tol"])
    if not np.isfinite(kt) or kt <= 0.0:
        raise ValueError(f"krylov_tol must be a finite positive float, got {kt!r}.")
    self.krylov_tol = kt


def _tdvp_settings(self, tdvp_sweeps, tdvp_mode) -> None:
    if isinstance(tdvp_sweeps, bool) or not isinstance(tdvp_sweeps, int):
        raise TypeError(f"tdvp_sweeps must be int, got {type(tdvp_sweeps).__name__}.")
    if tdvp_sweeps < 1:
        raise ValueError(f"tdvp_sweeps must be >= 1, got {tdvp_sweeps}.")
    self.tdvp_sweeps = tdvp_sweeps
    if tdvp_mode not in ("1site", "2site", "dynamic"):
        raise ValueError(f"tdvp_mode must be one of ('1site', '2site', 'dynamic'), got {tdvp_mode!r}.")
    self.tdvp_mode = tdvp_mode


class AnalogSimParams(_Ordering):
    """Hamiltonian-evolution parameters (:477-613).  Defaults: order=1, tdvp_mode="2site", preset "balanced".  Arguments are checked in
    the reference's order, with its exception types and messages (:579-613)."""

    def __init__(self, observables=None, elapsed_time: float = 0.1, dt: float = 0.1, num_traj: int | None = None,
                 max_bond_dim=_USE_PRESET, trunc_mode: str = "discarded_weight", svd_threshold: float | None = None,
                 krylov_tol: float | None = None, order: int = 1, *, preset: str = "balanced", sample_timesteps: bool = True,
                 evolution_mode: str = "tdvp", get_state: bool = False, random_seed: int | None = None,
                 multi_time_observables=None, tdvp_sweeps: int = 1, tdvp_mode: str = "2site") -> None:
        _validate_seed(random_seed)
        pv = _preset_values(preset)
        self.preset = preset
        self.observables = [] if observables is None else list(observables)
        _check_observable_mix(self.observables)
        n_steps = _time_grid(elapsed_time, dt)
        self.elapsed_time, self.dt = float(elapsed_time), float(dt)
        self.times = self.dt * np.arange(n_steps + 1, dtype=np.float64)
        if n_steps > 0:
            self.times[-1] = self.elapsed_time
        self.sample_timesteps = sample_timesteps
        _numerics(self, pv, num_traj, max_bond_dim, trunc_mode, svd_threshold, krylov_tol)
        self.order = order
        mode = getattr(evolution_mode, "value", evolution_mode)
        if mode not in ("tdvp", "bug"):
            raise ValueError(f"evolution_mode must be one of ('tdvp', 'bug'), got {evolution_mode!r}.")
        if mode != "tdvp":
            raise NotImplementedError("evolution_mode=BUG is outside the B200 hot-path scope (SURVEY.md 2.1 #20)")
        self.evolution_mode = "tdvp"
        self.get_state = get_state
        self.random_seed = random_seed
        self.multi_time_observables = [] if multi_time_observables is None else list(multi_time_observables)
        _tdvp_settings(self, tdvp_sweeps, tdvp_mode)


class DigitalSimParams(_Ordering):
    """Circuit-simulation parameters, keyword-only as in the reference (:616-750), checked in its order (:717-756)."""

    def __init__(self, *, observables=None, shots: int | None = None, num_traj: int | None = None, max_bond_dim=_USE_PRESET,
                 trunc_mode: str = "discarded_weight", svd_threshold: float | None = None, krylov_tol: float | None = None,
                 preset: str = "balanced", get_state: bool = False, sample_layers: bool = False, num_mid_measurements: int = 0,
                 random_seed: int | None = None, gate_mode: str = "mpo", tdvp_sweeps: int = 1, tdvp_mode: str = "2site") -> None:
        _validate_seed(random_seed)
        pv = _preset_values(preset)
        self.preset = preset
        self.observables = [] if observables is None else list(observables)
        _check_observable_mix(self.observables)
        if shots is not None and (isinstance(shots, bool) or not isinstance(shots, int) or shots < 1):
            raise ValueError(f"shots must be a positive int or None, got {shots!r}.")
        self.shots = shots
        _numerics(self, pv, num_traj, max_bond_dim, trunc_mode, svd_threshold, krylov_tol)
        self.get_state = get_state
        self.sample_layers = sample_layers
        self.num_mid_measurements = num_mid_measurements      # (:682,746) the run takes the count from the compiled circuit, as simulator.py:1723-1741 does
        self.random_seed = random_seed
        if gate_mode not in ("tdvp", "full-tdvp", "swaps", "mpo"):
            raise ValueError(f"gate_mode must be one of ('tdvp', 'full-tdvp', 'swaps', 'mpo'), got {gate_mode!r}.")
        self.gate_mode = gate_mode
        _tdvp_settings(self, tdvp_sweeps, tdvp_mode)


StrongSimParams = DigitalSimParams   # pre-0.6 names used by BASELINE.json's north_star (UPGRADING.md:49-60)
WeakSimParams = DigitalSimParams
