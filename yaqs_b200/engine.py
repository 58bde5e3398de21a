"""Batched Tensor Jump Method engine: all trajectories of a batch advance in lock-step on one GPU.

Host-side mirror of the reference's per-trajectory drivers (``analog/analog_tjm.py``) and methods
(``core/methods/tdvp``, ``dissipation.py``, ``stochastic_process.py``, the gauge moves of ``data_structures/mps.py``),
re-designed for a batch: every step below is ONE kernel launch over all trajectories (ragged bond dimensions live in a
device int32 array, data-dependent work is gated by device masks), and the only host synchronisations are the
"any trajectory still iterating?" reads of the Lanczos and Jacobi loops and the "did any trajectory jump?" read.

The arithmetic is done by ``kernels`` -- the ctypes binding of libyaqs_b200.so in production.  Tests may inject the
numpy twin from ``oracle/`` to exercise this host logic on CPU; the package itself never falls back to it.
"""
from __future__ import annotations

import copy
import os
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np
import scipy.linalg
import torch

from .gates import GateLibrary
from .rng import STREAM_SAMPLE, STREAM_TRAJECTORY
from .kernels import (DIST, LANCZOS_PITCH, OP_ABSORB_L, OP_ABSORB_R, OP_ENVL1, OP_ENVL3, OP_ENVR3, OP_EVEC, OP_GRAM, OP_LCONTRACT, OP_MERGE,
                      OP_RCONTRACT, OP_ROTATE, OP_ROTATE_TOP, SVD_COLS, SVD_META, SVD_ORIENT_COLS, SVD_ORIENT_ROWS, SVD_ROWS, SVD_THETA, TRUNC_MODES, Buf, Ext)

C128 = torch.complex128
KRYLOV_MAX = 25          # max_lanczos_iterations (matrix_exponential.py:37)
SHIFT_THRESHOLD = 1e-12  # hard-coded policy of SVD gauge shifts (mps.py:757-759)

# status bits reported per trajectory
STATUS_SVD_NOT_CONVERGED = 1
STATUS_NONFINITE = 2
STATUS_ZERO_JUMP_WEIGHT = 4
STATUS_IMAG_PART = 8            # an observable came out with imag >= 1e-13 (the reference asserts, mps.py:1233)


@dataclass
class EngineParams:
    """The fields of AnalogSimParams / DigitalSimParams the hot path reads (simulation_parameters.py:477-750)."""
    dt: float = 0.1
    n_steps: int = 10
    order: int = 2
    max_bond_dim: int | None = 128
    svd_threshold: float = 1e-6
    trunc_mode: str = "discarded_weight"
    krylov_tol: float = 1e-4
    tdvp_mode: str = "2site"
    tdvp_sweeps: int = 1
    sample_timesteps: bool = True
    digital: bool = False

    def __post_init__(self) -> None:
        # The reference stores any int as max_bond_dim (simulation_parameters.py:308-327) and fails inside the first split when it is
        # not positive; the batch sizes its slots from the cap, so it is refused here, where every run passes through.
        if self.max_bond_dim is not None and int(self.max_bond_dim) < 1:
            raise ValueError(f"max_bond_dim must be a positive int or None, got {self.max_bond_dim!r}.")

    @property
    def n_times(self) -> int:
        return self.n_steps + 1

    @property
    def tdvp_min_keep(self) -> int:
        """sweep_utils.py:33-46."""
        return 2 if self.max_bond_dim is None else min(2, self.max_bond_dim)


@dataclass
class JumpProcess:
    """One entry of the site-ordered jump table (stochastic_process.py:139-176)."""
    kind: str                       # "one" | "pauli2" | "adjacent"
    sites: tuple[int, ...]
    strength: float
    ldl: np.ndarray | None = None   # L^dagger L (2x2 or 4x4) for non-scalar weights
    ops: tuple = ()                 # jump operator(s): (L,) for one/adjacent, (F0, F1) for long-range Pauli


@dataclass
class Uniforms:
    """Where the two uniforms of one jump test come from (yb_jump_test): rows of a device table (streams replayed from the reference's
    seeded numpy generators, or pre-filled Philox tables) or the Philox generator inside the kernel; ``pos`` = per-trajectory stream
    position (advanced by the kernel by 1 + jumped), else column ``col`` for every trajectory."""
    table: torch.Tensor | None = None       # [B, n] float64, contiguous rows
    col: int = 0
    pos: torch.Tensor | None = None         # int64 [B]
    philox: tuple | None = None             # (seed, global trajectory indices int64 [B], stream tag, first block)


def layout_chi(params: "EngineParams") -> int | None:
    """Slot capacity the layout is built for: ``max_bond_dim``, except for tdvp_mode="dynamic", whose uncapped splits let a bond
    exceed the cap between two ``_cap_bonds`` passes (integrators.py:294-511): at most d * d * cap (a split is bounded by d times
    either neighbour bond, each of which may itself come from an uncapped split of capped bonds)."""
    if params.max_bond_dim is None or params.tdvp_mode != "dynamic":
        return params.max_bond_dim
    return 4 * params.max_bond_dim


class BatchLayout:
    """Slot capacities and offsets of the per-trajectory tensors (DESIGN.md "Data layout in HBM")."""

    def __init__(self, length: int, nbatch: int, chi_cap: int | None, mpo_bonds: Sequence[int], init_bonds: Sequence[int], d: int = 2,
                 bond_factor: Sequence[int] | None = None):
        """``bond_factor``: per-bond multiplier of the slot capacity (second capacity of the gate-MPO route: between
        ``MPO.multiply`` and the compression sweep a bond inside the gate support holds chi * D_gate, mpo.py:1534-1549)."""
        self.L, self.B, self.d = length, nbatch, d
        natural = [d ** min(j, length - j) for j in range(length + 1)]
        if chi_cap is None:
            if max(natural) > 4096:
                raise ValueError("max_bond_dim=None needs bonds up to d**(L/2); give an explicit max_bond_dim for this chain length")
            cap = natural
        else:
            cap = [min(chi_cap, n) for n in natural]
        self.cap = [max(c, int(b)) for c, b in zip(cap, init_bonds)]
        if bond_factor is not None:
            self.cap = [c * max(1, int(f)) for c, f in zip(self.cap, bond_factor)]
        self.cap[0] = self.cap[length] = 1
        self.D = [int(x) for x in mpo_bonds]                       # MPO bond dims, D[0] = D[L] = 1 for open chains
        self.slot = [d * self.cap[i] * self.cap[i + 1] for i in range(length)]
        self.site_off = [0] * length
        acc = 0
        for i in range(length):
            self.site_off[i] = acc
            acc += self.slot[i] * nbatch
        self.state_elems = acc
        self.env_slot = [self.cap[j] * self.D[j] * self.cap[j] for j in range(length + 1)]
        self.env_off = [0] * (length + 1)
        acc = 0
        for j in range(length + 1):
            self.env_off[j] = acc
            acc += self.env_slot[j] * nbatch
        self.env_elems = acc
        self.cmax = max(self.cap)
        self.Dmax = max(self.D)
        pair = [self.cap[i] * self.cap[i + 2] for i in range(length - 1)] or [1]
        single = [self.cap[i] * self.cap[i + 1] for i in range(length)]
        self.theta_slot = d * d * max(pair)
        self.nloc_max = max(self.theta_slot, d * max(single))
        self.t_slot = self.nloc_max * self.Dmax
        self.svd_dim = d * self.cmax


class BatchedState:
    """B matrix-product states: packed site tensors + the device-resident bond dimensions (mps.py:54-69)."""

    def __init__(self, layout: BatchLayout, device):
        self.layout = layout
        self.A = torch.zeros(layout.state_elems, dtype=C128, device=device)
        self.bond = torch.ones((layout.B, layout.L + 1), dtype=torch.int32, device=device)

    def site(self, i: int) -> Buf:
        return Buf(self.A, self.layout.site_off[i], self.layout.slot[i])

    def ext(self, j: int) -> Ext:
        return Ext(self.bond, j, self.layout.L + 1)

    def copy_from(self, other: "BatchedState") -> None:
        self.A.copy_(other.A)
        self.bond.copy_(other.bond)

    def transfer_from(self, other: "BatchedState") -> None:
        """Device-to-device hand-off between two engines whose slot capacities differ (SimulationProgram segments with different
        ``max_bond_dim``): site tensors are stored compactly (d, l, r) at the start of their slot, so the leading
        ``min(slot_dst, slot_src)`` elements of every slot carry the whole tensor as long as the destination capacities cover the
        source's current bonds (BatchedTJM sizes its layout from them)."""
        lay, src = self.layout, other.layout
        if (lay.L, lay.B, lay.d) != (src.L, src.B, src.d):
            raise ValueError("transfer_from needs the same chain length, batch size and physical dimension")
        top = other.bond.max(dim=0).values.cpu().tolist()
        if any(int(t) > c for t, c in zip(top, lay.cap)):
            raise ValueError(f"bonds {top} exceed the destination capacities {lay.cap}")
        self.bond.copy_(other.bond)
        for i in range(lay.L):
            n = min(lay.slot[i], src.slot[i])
            dst = self.A[lay.site_off[i]: lay.site_off[i] + lay.B * lay.slot[i]].view(lay.B, lay.slot[i])
            dst[:, :n].copy_(other.A[src.site_off[i]: src.site_off[i] + src.B * src.slot[i]].view(src.B, src.slot[i])[:, :n])

    def load(self, tensors_per_traj: Sequence[Sequence[np.ndarray]] | Sequence[np.ndarray]) -> None:
        """Upload host MPS tensors (s, l, r).  A single MPS (list of arrays) is uploaded once and broadcast to all trajectories on
        the device; per-trajectory states are staged site by site (bounded host memory: one site of the batch at a time)."""
        lay = self.layout
        single = isinstance(tensors_per_traj[0], np.ndarray)
        bond = torch.ones((lay.B, lay.L + 1), dtype=torch.int32)
        dev = self.A.device
        for i in range(lay.L):
            per = [tensors_per_traj[i]] if single else [tensors_per_traj[b][i] for b in range(lay.B)]
            stage = torch.zeros((len(per), lay.slot[i]), dtype=C128)
            sn = stage.numpy()
            for row, t in enumerate(per):
                d, l, r = t.shape
                if l > lay.cap[i] or r > lay.cap[i + 1] or d != lay.d:
                    raise ValueError(f"site {i}: tensor {t.shape} exceeds the slot capacity ({lay.d},{lay.cap[i]},{lay.cap[i + 1]})")
                sn[row, : t.size] = np.ascontiguousarray(t, dtype=np.complex128).reshape(-1)
                if single:
                    bond[:, i], bond[:, i + 1] = l, r
                else:
                    bond[row, i], bond[row, i + 1] = l, r
            dst = self.A[lay.site_off[i]: lay.site_off[i] + lay.B * lay.slot[i]].view(lay.B, lay.slot[i])
            dst.copy_(stage.to(dev).expand(lay.B, lay.slot[i]) if single else stage.to(dev))
        self.bond.copy_(bond)

    def download(self, b: int) -> list[np.ndarray]:
        lay = self.layout
        bond = self.bond[b].cpu().numpy()
        out = []
        for i in range(lay.L):
            l, r = int(bond[i]), int(bond[i + 1])
            off = lay.site_off[i] + b * lay.slot[i]
            out.append(self.A[off: off + lay.d * l * r].cpu().numpy().reshape(lay.d, l, r).copy())
        return out


class BatchedTJM:
    """Lock-step TJM over a batch.  ``kernels`` is a CudaKernels instance (production) or a test double."""

    def __init__(self, kernels, length: int, nbatch: int, mpo: Sequence[np.ndarray], params: EngineParams,
                 noise: Sequence[dict] | None, observables: Sequence[tuple[np.ndarray, Sequence[int]]],
                 init_tensors, svd_block_rows: int | None = None, gram_min_k: int = 33, scheduled_jumps: Sequence[dict] | None = None,
                 extra_mpos: Sequence[Sequence[np.ndarray]] = (), bond_factor: Sequence[int] | None = None):
        """``init_tensors``: one MPS (list of (d, l, r) arrays, broadcast), one MPS per trajectory, or the BatchedState of another
        engine (device-to-device hand-off between SimulationProgram segments).  ``extra_mpos``: further MPOs this engine will be
        switched to (``run(interval_mpos=...)``); the layout takes the per-bond maximum of their bond dimensions and every MPO is
        zero-padded to it on upload."""
        self.K = kernels
        self._const_cache = {}
        self.scheduled_jumps = list(scheduled_jumps) if scheduled_jumps else []       # normalised {time, sites, matrix} (order 1 only)
        self.dev = kernels.device
        self.p = params
        self.L, self.B, self.d = length, nbatch, 2
        from_state = isinstance(init_tensors, BatchedState)
        init_bonds = [1] * (length + 1)
        if from_state:
            if init_tensors.layout.B != nbatch or init_tensors.layout.L != length:
                raise ValueError("the handed-over BatchedState has another batch size or chain length")
            init_bonds = [int(v) for v in init_tensors.bond.max(dim=0).values.cpu().tolist()]
        else:
            single = isinstance(init_tensors[0], np.ndarray)
            for tensors in ([init_tensors] if single else init_tensors):
                for i, t in enumerate(tensors):
                    init_bonds[i] = max(init_bonds[i], t.shape[1])
                    init_bonds[i + 1] = max(init_bonds[i + 1], t.shape[2])
        mpo_bonds = [int(mpo[0].shape[2])] + [int(w.shape[3]) for w in mpo]
        for other in extra_mpos:
            ob = [int(other[0].shape[2])] + [int(w.shape[3]) for w in other]
            mpo_bonds = [max(a, b) for a, b in zip(mpo_bonds, ob)]
        self.lay = BatchLayout(length, nbatch, layout_chi(params), mpo_bonds, init_bonds, bond_factor=bond_factor)
        lay = self.lay
        dev = self.dev
        self.phi = BatchedState(lay, dev)
        self.psi = BatchedState(lay, dev) if params.order == 2 else None
        if from_state:
            self.phi.transfer_from(init_tensors)
        else:
            self.phi.load(init_tensors)
        # environments: bond-slot storage; the left and right blocks of a 2-site sweep never need the same bond at the
        # same time, 1TDVP's bond update does.
        self.renv = torch.zeros(lay.env_elems, dtype=C128, device=dev)
        self.lenv = torch.zeros(lay.env_elems, dtype=C128, device=dev) if params.tdvp_mode in ("1site", "dynamic") else self.renv
        # MPO tensors (shared by all trajectories): W[o,p,a,a'], its left-environment transpose WT[p,o,a',a], merged pairs
        self._mpo_cache: dict = {}
        self._set_mpo(self._upload_mpo(mpo))
        # workspaces
        B = nbatch
        self.theta = torch.zeros(B * lay.theta_slot, dtype=C128, device=dev)
        self.T1 = torch.zeros(B * lay.t_slot, dtype=C128, device=dev)
        self.T2 = torch.zeros(B * lay.t_slot, dtype=C128, device=dev)
        self.svec = lay.nloc_max
        self.basis = torch.zeros(B * (KRYLOV_MAX + 1) * self.svec, dtype=C128, device=dev)
        self.tmp = torch.zeros(B * lay.nloc_max, dtype=C128, device=dev)
        sd = lay.svd_dim
        self.svd_w = torch.zeros(B * sd * sd, dtype=C128, device=dev)
        self.svd_q = torch.zeros(B * sd * sd, dtype=C128, device=dev)
        self.svd_u = torch.zeros(B * sd * sd, dtype=C128, device=dev)
        self.svd_v = torch.zeros(B * sd * sd, dtype=C128, device=dev)
        self.svd_s = torch.zeros(B * sd, dtype=torch.float64, device=dev)
        self.svd_meta = torch.zeros(B * SVD_META, dtype=torch.int32, device=dev)
        self.svd_rot = torch.zeros(B, dtype=torch.int32, device=dev)
        self.keep = torch.ones(B, dtype=torch.int32, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.lz_count = torch.zeros(2, dtype=torch.int32, device=dev)
        # jump test / selection outputs (yb_jump_test, yb_jump_select)
        self.jt_dp = torch.zeros(B, dtype=torch.float64, device=dev)
        self.jt_u1 = torch.zeros(B, dtype=torch.float64, device=dev)
        self.jt_u2 = torch.zeros(B, dtype=torch.float64, device=dev)
        self.jt_mask = torch.zeros(B, dtype=torch.int32, device=dev)
        self.jt_n = torch.zeros(1, dtype=torch.int32, device=dev)
        self.jt_total = torch.zeros(B, dtype=torch.float64, device=dev)
        self.jt_choice = torch.zeros(B, dtype=torch.int32, device=dev)
        self.jt_ops = torch.zeros(B * self.JT_NOPS, dtype=C128, device=dev)
        self.use_graphs = os.environ.get("YB_GRAPHS", "1") != "0" and bool(getattr(kernels, "supports_graphs", False))
        self._graphs: dict = {}
        self._graph_seen: set = set()
        self.g_mask = torch.zeros(B, dtype=torch.int32, device=dev)
        self.g_work = torch.zeros(B, dtype=torch.int32, device=dev)
        self.jt_w = None                                    # [B, n_proc] weights / pdf of the current noise context (allocated per context)
        self.philox = None                                  # (seed, global trajectory indices int64 [B], sample offset): draw the uniforms in the kernel
        self.status = torch.zeros(B, dtype=torch.int32, device=dev)
        self.alpha = torch.zeros(B * LANCZOS_PITCH, dtype=torch.float64, device=dev)
        self.beta = torch.zeros(B * LANCZOS_PITCH, dtype=torch.float64, device=dev)
        self.coef = torch.zeros(B * LANCZOS_PITCH, dtype=C128, device=dev)
        self.nrm = torch.zeros(B, dtype=torch.float64, device=dev)
        self.active = torch.zeros(B, dtype=torch.int32, device=dev)
        self.kdone = torch.zeros(B, dtype=torch.int32, device=dev)
        self.norm_buf = torch.zeros(B, dtype=torch.float64, device=dev)
        self.scale_buf = torch.zeros(B, dtype=torch.float64, device=dev)
        self.rdm1 = torch.zeros(B * 4, dtype=C128, device=dev)
        self._rdm1_all = self._rdm2_all = None          # [L, B, 4] / [L - 1, B, 16] tables of rho_sweep (allocated on first use)
        self._obs_tab = None
        self._spectra_now: dict = {}
        # Jacobi panel geometry is chosen per call from the capacities of the matrix at hand (_jacobi_geometry); a test may force
        # the panel height to exercise the multi-block schedule on small matrices
        self.svd_block_rows_forced = svd_block_rows
        # Gram-matrix SVD route (Hermitian eigen-solver) for matrices whose Jacobi side would exceed one shared-memory panel
        self.gram_min_k = gram_min_k
        self._jac_geo: dict = {}
        # project_site / update_right_environment: the MPO-tensor step runs in the epilogue of the first GEMM (yb_rcontract_mpo);
        # YB_FUSED_MPO=0 selects the separate k_gemm + k_wapply launches (A/B runs)
        self.fused_mpo = os.environ.get("YB_FUSED_MPO", "1") != "0"
        # Gram-route eigen-solver: yb_heev_rows (blocked, round 2) unless YB_HEEV_ROWS=0 selects the round-1 kernels + EVEC product
        self.heev_rows = os.environ.get("YB_HEEV_ROWS", "1") != "0"
        # keep-all gauge shifts by Cholesky-QR (round 2, second session); YB_CHOLQR=0 sends them through the eigen-solver (A/B runs)
        self.cholqr = os.environ.get("YB_CHOLQR", "1") != "0"
        # capped splits back-transform only the eigenvector rows they can keep (yb_heev_rows_top); YB_HEEV_TOP=0 for A/B runs
        self.heev_top = os.environ.get("YB_HEEV_TOP", "1") != "0"
        self.use_gram = sd >= gram_min_k and sd <= 512
        if self.use_gram:
            self.heev_g = torch.zeros(B * sd * sd, dtype=C128, device=dev)
            self.heev_q = torch.zeros(B * sd * sd, dtype=C128, device=dev)
            self.heev_z = torch.zeros(B * sd * sd, dtype=C128, device=dev)
            self.heev_zr = torch.zeros(B * sd * sd, dtype=torch.float64, device=dev)
            self.heev_d = torch.zeros(B * sd, dtype=torch.float64, device=dev)
            self.heev_e = torch.zeros(B * sd, dtype=torch.float64, device=dev)
            self.heev_tau = torch.zeros(B * sd, dtype=C128, device=dev)
            self.heev_s = torch.zeros(B * sd * sd, dtype=C128, device=dev)
            self.sel_n = torch.zeros(B, dtype=torch.int32, device=dev)
            self.sel_idx = torch.zeros(B * sd, dtype=torch.int32, device=dev)
            self.status_scratch = torch.zeros(B, dtype=torch.int32, device=dev)
            self.amb = torch.zeros(B, dtype=torch.int32, device=dev)
            self.heev_order = torch.zeros(B * sd, dtype=torch.int32, device=dev)
            self.heev_ds = torch.zeros(B * sd, dtype=torch.float64, device=dev)
            self.heev_ntop = torch.zeros(B, dtype=torch.int32, device=dev)
        # noise tables
        self._noise_cache: dict = {}
        self._set_noise(self._make_noise_ctx(noise))
        # observables (already in the reference's sorted order)
        # entries: (matrix, sites) | ("entropy", (i, i+1)) | ("pvm:<bitstring>", ())
        self.observables = [(m if isinstance(m, str) else np.asarray(m, dtype=np.complex128), tuple(int(s) for s in sites)) for m, sites in observables]
        self._ones_i32 = torch.ones(B, dtype=torch.int32, device=dev)
        self.on_sample = None                               # callback(engine, column, results, diagnostics) after every recorded column of run_digital
        self.stats = {"krylov_calls": 0, "matvecs": 0, "matvecs_2site": 0, "svd_calls": 0, "svd_sweeps": 0, "jumps": 0, "gram_svd_calls": 0, "gram_fallbacks": 0, "cholqr_calls": 0,
                      "cholqr_fallbacks": 0, "heev_top_calls": 0, "graphs_captured": 0, "graph_launches": 0, "graph_capture_failed": 0}
        self.graph_error = None

    # ------------------------------------------------------------------------------------------------------------------
    # operator / noise contexts (one each on the analog path; per gate on the digital path)
    # ------------------------------------------------------------------------------------------------------------------
    def _upload_mpo(self, mpo: Sequence[np.ndarray]) -> dict:
        dev = self.dev
        D = self.lay.D
        if any(w.shape[2] != D[i] or w.shape[3] != D[i + 1] for i, w in enumerate(mpo)):      # zero-pad to the layout's bond dims
            padded = []
            for i, w in enumerate(mpo):
                if w.shape[2] > D[i] or w.shape[3] > D[i + 1]:
                    raise ValueError(f"MPO site {i}: bonds {w.shape[2:]} exceed the layout's ({D[i]}, {D[i + 1]}); pass it in extra_mpos")
                z = np.zeros((w.shape[0], w.shape[1], D[i], D[i + 1]), dtype=np.complex128)
                z[:, :, : w.shape[2], : w.shape[3]] = w
                padded.append(z)
            mpo = padded
        w1 = [torch.from_numpy(np.ascontiguousarray(w, dtype=np.complex128).reshape(-1)).to(dev) for w in mpo]
        w1t = [torch.from_numpy(np.ascontiguousarray(np.transpose(w, (1, 0, 3, 2)), dtype=np.complex128).reshape(-1)).to(dev) for w in mpo]
        w2 = []
        for i in range(len(mpo) - 1):
            m = np.einsum("acei,bdif->abcdef", mpo[i], mpo[i + 1])
            sh = m.shape
            w2.append(torch.from_numpy(np.ascontiguousarray(m.reshape(sh[0] * sh[1], sh[2] * sh[3], sh[4], sh[5])).reshape(-1)).to(dev))
        return dict(W1=w1, W1T=w1t, W2=w2)

    def _set_mpo(self, ms: dict) -> None:
        self.W1, self.W1T, self.W2 = ms["W1"], ms["W1T"], ms["W2"]

    def _make_noise_ctx(self, noise) -> dict:
        ctx = dict(noise=list(noise) if (noise or isinstance(noise, list) and self.p.digital) else None, diss={}, table=[])
        ctx["active"] = ctx["noise"] is not None and not all(pr["strength"] == 0 for pr in ctx["noise"])
        return ctx

    def _set_noise(self, ctx: dict) -> None:
        self._nctx = ctx
        self.noise, self.noise_active, self._diss_cache = ctx["noise"], ctx["active"], ctx["diss"]
        if self.noise is not None and not ctx["table"]:
            ctx["table"] = self._build_jump_table()
        self.jump_table = ctx["table"]

    # ------------------------------------------------------------------------------------------------------------------
    # small helpers
    # ------------------------------------------------------------------------------------------------------------------
    JT_NOPS = 24                 # operator row of a process: one-site op on its first site (4), on its second site (4), two-site op (16)

    def env(self, j: int, left: bool) -> Buf:
        return Buf(self.lenv if left else self.renv, self.lay.env_off[j], self.lay.env_slot[j])

    def _buf(self, t: torch.Tensor, slot: int) -> Buf:
        return Buf(t, 0, slot)

    def _const(self, array) -> torch.Tensor:
        """Device copy of a small host constant (gate, dissipator, jump operator), uploaded once and cached by value: the same
        2 x 2 / 4 x 4 operators recur at every site, step and gate, and a pageable upload per use is a host synchronisation."""
        a = np.ascontiguousarray(array, dtype=np.complex128).reshape(-1)
        key = a.tobytes()
        t = self._const_cache.get(key)
        if t is None:
            t = self._const_cache[key] = torch.from_numpy(a.copy()).to(self.dev)
        return t

    def _set_bond(self, st: BatchedState, j: int, mask) -> None:
        if mask is None:
            st.bond[:, j] = self.keep
        else:
            st.bond[:, j] = torch.where(mask != 0, self.keep, st.bond[:, j])

    def _trunc_args(self, mode: str, threshold: float, max_bond: int | None, min_keep: int):
        return TRUNC_MODES[mode], float(threshold), int(max_bond) if max_bond else 0, int(min_keep)

    # ------------------------------------------------------------------------------------------------------------------
    # contractions
    # ------------------------------------------------------------------------------------------------------------------
    def heff_matvec(self, st: BatchedState, i: int, nsites: int, x: Buf, y: Buf, mask) -> None:
        """project_site (primitives.py:180-204) on sites i..i+nsites-1: three launches."""
        lay, K, B = self.lay, self.K, self.B
        dphys = self.d ** nsites
        Dl, Dr = lay.D[i], lay.D[i + nsites]
        el, er = st.ext(i), st.ext(i + nsites)
        cl, cr = lay.cap[i], lay.cap[i + nsites]
        w = self.W1[i] if nsites == 1 else self.W2[i]
        t1, t2 = self._buf(self.T1, lay.t_slot), self._buf(self.T2, lay.t_slot)
        if self.fused_mpo and Dl <= 8 and Dr <= 8 and (dphys * cl > 64 or Dr * cr > 64):      # small products: thread-per-element kernels instead
            K.rcontract_mpo(B, mask, el, er, dphys, Dl, Dr, x, self.env(i + nsites, False), w, t2, cl, cr)
        else:
            K.gemm(OP_RCONTRACT, B, 1, mask, el, er, None, dphys, Dl, Dr, x, self.env(i + nsites, False), t1, 0, dphys * cl, Dr * cr)
            K.wapply(B, mask, el, er, dphys, dphys, Dl, Dr, w, t1, t2, cl * cr)
        K.gemm(OP_LCONTRACT, B, dphys, mask, el, er, None, dphys, Dl, Dr, self.env(i, True), t2, y, 0, cl, cr)
        self.stats["matvecs"] += 1
        self.stats["matvecs_2site"] += int(nsites == 2)

    def bond_matvec(self, el: Ext, er: Ext, jbond: int, x: Buf, y: Buf, mask, cl: int, cr: int) -> None:
        """project_bond (primitives.py:207-226) at bond jbond."""
        lay, K, B = self.lay, self.K, self.B
        D = lay.D[jbond]
        t1 = self._buf(self.T1, lay.t_slot)
        K.gemm(OP_RCONTRACT, B, 1, mask, el, er, None, 1, D, D, x, self.env(jbond, False), t1, 0, cl, D * cr)
        K.gemm(OP_LCONTRACT, B, 1, mask, el, er, None, 1, D, D, self.env(jbond, True), t1, y, 0, cl, cr)
        self.stats["matvecs"] += 1

    def env_boundaries(self, st: BatchedState) -> None:
        """Identity blocks at both chain ends (primitives.py:162-168, integrators.py:186-193)."""
        lay = self.lay
        self.K.env_identity(self.B, None, st.ext(lay.L), lay.D[lay.L], self.env(lay.L, False), lay.cap[lay.L])
        self.K.env_identity(self.B, None, st.ext(0), lay.D[0], self.env(0, True), lay.cap[0])

    def env_update_right(self, st: BatchedState, s: int) -> None:
        """update_right_environment (primitives.py:77-108): block at bond s+1 -> block at bond s."""
        lay, K, B, d = self.lay, self.K, self.B, self.d
        Dl, Dr = lay.D[s], lay.D[s + 1]
        el, er = st.ext(s), st.ext(s + 1)
        cl, cr = lay.cap[s], lay.cap[s + 1]
        a = st.site(s)
        t1, t2 = self._buf(self.T1, lay.t_slot), self._buf(self.T2, lay.t_slot)
        if self.fused_mpo and Dl <= 8 and Dr <= 8 and (d * cl > 64 or Dr * cr > 64):
            K.rcontract_mpo(B, None, el, er, d, Dl, Dr, a, self.env(s + 1, False), self.W1[s], t2, cl, cr)
        else:
            K.gemm(OP_RCONTRACT, B, 1, None, el, er, None, d, Dl, Dr, a, self.env(s + 1, False), t1, 0, d * cl, Dr * cr)
            K.wapply(B, None, el, er, d, d, Dl, Dr, self.W1[s], t1, t2, cl * cr)
        K.gemm(OP_ENVR3, B, 1, None, el, er, None, d, Dl, Dr, t2, a, self.env(s, False), 0, Dl * cl, cl)

    def env_update_left(self, st: BatchedState, s: int) -> None:
        """update_left_environment (primitives.py:111-136): block at bond s -> block at bond s+1."""
        lay, K, B, d = self.lay, self.K, self.B, self.d
        Dl, Dr = lay.D[s], lay.D[s + 1]
        el, er = st.ext(s), st.ext(s + 1)
        cl, cr = lay.cap[s], lay.cap[s + 1]
        a = st.site(s)
        t1, t2 = self._buf(self.T1, lay.t_slot), self._buf(self.T2, lay.t_slot)
        K.gemm(OP_ENVL1, B, d, None, el, er, None, d, Dl, Dr, self.env(s, True), a, t1, 0, Dl * cl, cr)
        K.wapply(B, None, el, er, d, d, Dr, Dl, self.W1T[s], t1, t2, cl * cr)
        K.gemm(OP_ENVL3, B, 1, None, el, er, None, d, Dl, Dr, a, t2, self.env(s + 1, True), 0, cr, Dr * cr)

    def init_right_envs(self, st: BatchedState, lo: int = 0, hi: int | None = None) -> None:
        """initialize_right_environments (primitives.py:139-174) for the sites lo..hi (a gate window, or the chain)."""
        hi = self.L - 1 if hi is None else hi
        lay = self.lay
        self.K.env_identity(self.B, None, st.ext(hi + 1), lay.D[hi + 1], self.env(hi + 1, False), lay.cap[hi + 1])
        self.K.env_identity(self.B, None, st.ext(lo), lay.D[lo], self.env(lo, True), lay.cap[lo])
        for s in reversed(range(lo + 1, hi + 1)):
            self.env_update_right(st, s)

    def merge(self, st: BatchedState, i: int, mask=None) -> Buf:
        """merge_two_site (decompositions.py:87-102) of sites i, i+1 into the theta workspace."""
        lay = self.lay
        th = self._buf(self.theta, lay.theta_slot)
        self.K.gemm(OP_MERGE, self.B, self.d * self.d, mask, st.ext(i), st.ext(i + 1), st.ext(i + 2), self.d, 1, 1, st.site(i), st.site(i + 1),
                    th, 0, lay.cap[i], lay.cap[i + 2])
        return th

    # ------------------------------------------------------------------------------------------------------------------
    # Lanczos exponential (matrix_exponential.py:33-201)
    # ------------------------------------------------------------------------------------------------------------------
    def krylov(self, x: Buf, out: Buf, el: Ext, er: Ext, dphys: int, matvec, dt: float, mask) -> None:
        K, B = self.K, self.B
        basis = self._buf(self.basis, (KRYLOV_MAX + 1) * self.svec)
        # "anyone still iterating?" lives in a pair of device counters: step j counts into c[j & 1] and clears c[(j + 1) & 1] in the
        # same launch (lanczos_init clears c[0]), so the iteration issues no fill launch between its steps
        c2 = self.lz_count
        K.lanczos_init(B, mask, el, er, dphys, x, basis, self.svec, self.nrm, self.active, self.kdone, c2[0:1])
        self.stats["krylov_calls"] += 1
        for j in range(KRYLOV_MAX):
            matvec(basis.at(j * self.svec), basis.at((j + 1) * self.svec), self.active)
            cur, nxt = c2[(j & 1):(j & 1) + 1], c2[((j + 1) & 1):((j + 1) & 1) + 1]
            K.lanczos_step(B, j, KRYLOV_MAX, dt, self.p.krylov_tol, el, er, dphys, basis, self.svec, self.alpha, self.beta, self.nrm,
                           self.active, self.kdone, self.coef, cur, nxt)
            # convergence is per trajectory on the device (`active`); the host only needs to know when nobody is left.  The error
            # estimate needs j >= 1 and rarely passes before the third vector, so the first two iterations are not read back
            if j >= 2 and K.sync_int(cur) == 0:
                break
        K.lanczos_combine(B, mask, el, er, dphys, basis, self.svec, self.coef, KRYLOV_MAX, self.kdone, out)

    def update_site(self, st: BatchedState, i: int, nsites: int, x: Buf, dt: float, mask=None) -> None:
        """update_site (primitives.py:484-517): x <- exp(-i dt H_eff) x in place."""
        self.krylov(x, x, st.ext(i), st.ext(i + nsites), self.d ** nsites,
                    lambda src, dst, m: self.heff_matvec(st, i, nsites, src, dst, m), dt, mask)

    # ------------------------------------------------------------------------------------------------------------------
    # SVD driver + splits / gauge shifts
    # ------------------------------------------------------------------------------------------------------------------
    def _svd(self, mode: int, src: Buf, el: Ext, er: Ext, trunc, dist: str, mask, kcap: int = 0, dims: tuple[int, int] | None = None) -> None:
        """Truncated SVD of every trajectory's matrix.  kcap > 0 (capacity of the side that becomes the isometry) selects the
        Gram-matrix route for large matrices; the one-sided Jacobi is used otherwise and as its accurate fallback.  ``dims``:
        capacities (rows, cols) of the matrix -- they size the Jacobi panels (default: the layout's largest matrix)."""
        if self.use_gram and kcap > 0 and kcap >= self.gram_min_k and dist in ("left", "right") and trunc[0] in (0, 1):
            tm, thr, mb, _mk = trunc
            kb, lb = (0, 0) if dims is None else ((dims[1], dims[0]) if dist == "right" else (dims[0], dims[1]))      # bond side, other side
            if (self.cholqr and mode in (SVD_ROWS, SVD_COLS) and tm == 0 and thr <= self.CHOLQR_MAX_THRESHOLD and 0 < kb <= min(lb, self.CHOLQR_KMAX)
                    and (mb == 0 or mb >= kb)):
                self._svd_cholqr(mode, src, el, er, trunc, dist, mask, dims)
            else:
                self._svd_gram(mode, src, el, er, trunc, dist, mask, dims)
        else:
            self._svd_jacobi(mode, src, el, er, trunc, dist, mask, dims)

    # Gauge shifts whose truncation policy can only ever remove singular values below ~1e-6 (the hard-coded 1e-12 discarded weight of
    # mps.py:757-759, or none at all) keep every singular value unless the matrix is numerically rank-deficient.  For those the
    # factorisation is a Cholesky-QR with a keep-all certificate (csrc/yb_cholqr.cuh); trajectories the certificate does not cover
    # (rank-deficient bonds right after a lowering jump, a bond side longer than the other side) take the eigen-solver route.
    CHOLQR_MAX_THRESHOLD = 1e-9
    CHOLQR_KMAX = 128
    CHOLQR_TOL = 1e-7             # accepted orthonormality defect of the first pass; one Newton-Schulz step squares it

    def _svd_cholqr(self, mode: int, src: Buf, el: Ext, er: Ext, trunc, dist: str, mask, dims: tuple[int, int]) -> None:
        self._cholqr_chain(mode, src, el, er, trunc, dist, mask, dims)
        nfb = self.K.sync_int(self.counter)
        if nfb > 0:                                          # not certified: the eigen-solver route decides (and truncates) for those
            self.stats["cholqr_fallbacks"] += nfb
            self._svd_gram(mode, src, el, er, trunc, dist, self.amb.clone(), dims)

    def _cholqr_chain(self, mode: int, src: Buf, el: Ext, er: Ext, trunc, dist: str, mask, dims: tuple[int, int]) -> None:
        """The launches of one Cholesky-QR factorisation, nothing else: no host read, no host decision -- the sequence can be captured
        into a CUDA graph.  Leaves the factors of the accepted trajectories in svd_u / svd_v / keep, the flags in ``amb`` and their
        number in ``counter``."""
        K, B, lay = self.K, self.B, self.lay
        sd = lay.svd_dim
        sq = sd * sd
        w, linv = self._buf(self.svd_w, sq), self._buf(self.svd_q, sq)
        g, y1, g2, y2 = self._buf(self.heev_g, sq), self._buf(self.heev_q, sq), self._buf(self.heev_z, sq), self._buf(self.heev_s, sq)
        x, l2 = g, g2                                        # the Gram matrices are dead once factor / refine have read them
        # the bond side must index the rows of W: W = M^H for a right shift (U = site i becomes the isometry), W = M for a left shift
        orient = SVD_ORIENT_COLS if dist == "right" else SVD_ORIENT_ROWS
        kdim = max(1, min(sd, dims[1] if dist == "right" else dims[0]))
        ldim = max(1, min(sd, dims[0] if dist == "right" else dims[1]))
        K.svd_prepare(B, mask, mode | orient, el, er, self.d, src, w, linv, self.svd_meta, self.svd_rot, self.counter)      # (clears the flag counter)
        ek, eln = Ext(self.svd_meta, 0, SVD_META), Ext(self.svd_meta, 1, SVD_META)
        K.gemm(OP_GRAM, B, 1, mask, ek, eln, None, 1, 1, 1, w, w, g, 0, kdim, kdim)
        K.cholqr_factor(B, mask, self.svd_meta, g, linv, float(trunc[1]), self.amb, self.counter, kdim)
        K.gemm(OP_ROTATE, B, 1, mask, ek, eln, None, 1, 1, 1, linv, w, y1, 0, kdim, ldim)
        K.gemm(OP_GRAM, B, 1, mask, ek, eln, None, 1, 1, 1, y1, y1, g2, 0, kdim, kdim)
        K.cholqr_refine(B, mask, self.svd_meta, g2, x, self.CHOLQR_TOL, self.amb, self.counter)
        K.gemm(OP_ROTATE, B, 1, mask, ek, eln, None, 1, 1, 1, x, y1, y2, 0, kdim, ldim)
        K.gemm(OP_GRAM, B, 1, mask, ek, eln, None, 1, 1, 1, w, y2, l2, 0, kdim, kdim)
        K.cholqr_finalize(B, mask, self.svd_meta, y2, l2, Ext(self.keep, 0, 1), self._buf(self.svd_u, sq), self._buf(self.svd_v, sq), self.amb,
                          kdim * (ldim + kdim))

    def _cholqr_applies(self, trunc, dist: str, dims: tuple[int, int], kcap: int) -> bool:
        """The routing rule of ``_svd`` for gauge shifts: Gram route wanted, keep-all policy, bond side fits the shared-memory factorisation."""
        tm, thr, mb, _mk = trunc
        kb, lb = (dims[1], dims[0]) if dist == "right" else (dims[0], dims[1])
        return (self.use_gram and self.cholqr and kcap > 0 and kcap >= self.gram_min_k and tm == 0 and thr <= self.CHOLQR_MAX_THRESHOLD
                and 0 < kb <= min(lb, self.CHOLQR_KMAX) and (mb == 0 or mb >= kb))

    # ------------------------------------------------------------------------------------------------------------------
    # CUDA graphs of fixed launch sequences (include/yaqs_b200.h: yb_graph_*)
    # ------------------------------------------------------------------------------------------------------------------
    def _run_graphed(self, key, fn) -> None:
        """Run the launch sequence ``fn`` -- which must consist of library launches only, all decided by host-known parameters that
        ``key`` captures -- as a CUDA graph: the first use runs eagerly (kernels configure their attributes outside a capture), the
        second is captured, every later one is ONE host call.  Eager whenever graphs are off, the backend has none (numpy twin) or
        per-launch timing is active (bench.py's instrumented step)."""
        K = self.K
        if not self.use_graphs or K.timing_active or K.on_default_stream():
            fn()
            return
        handle = self._graphs.get(key)
        if handle is None:
            if key not in self._graph_seen:
                self._graph_seen.add(key)
                fn()
                return
            try:
                K.graph_begin()
                try:
                    fn()
                finally:
                    handle = K.graph_end()
            except Exception as exc:                         # capture refused: stay eager for the rest of the run
                self.use_graphs = False
                self.graph_error = repr(exc)
                self.stats["graph_capture_failed"] += 1
                fn()
                return
            self._graphs[key] = handle
            self.stats["graphs_captured"] += 1
        K.graph_launch(handle)
        self.stats["graph_launches"] += 1

    def release_graphs(self) -> None:
        for handle in self._graphs.values():
            self.K.graph_destroy(handle)
        self._graphs.clear()

    def __del__(self):
        try:
            self.release_graphs()
        except Exception:
            pass

    def _jacobi_geometry(self, dims: tuple[int, int] | None) -> tuple[int, int, int, int, int]:
        """(block_rows, nblocks, inner_sweeps, max_len, max_k) of the one-sided Jacobi for matrices of at most ``dims`` = (rows,
        cols): the rows being orthogonalised are those of the smaller side (k = min, length = max).  Panel height = the largest
        power of two <= 16 whose 2 * rows * (len + k) complex panel fits in ~200 KB of shared memory; matrices whose smaller side
        fits one panel pair (k <= 2 * rows: every chain-end site) take the single-launch schedule with up to 30 in-kernel sweeps."""
        sd = self.lay.svd_dim
        kmax, lmax = (sd, sd) if dims is None else (max(1, min(dims)), max(1, max(dims)))
        kmax, lmax = min(kmax, sd), min(lmax, sd)
        key = (kmax, lmax)
        geo = self._jac_geo.get(key)
        if geo is None:
            rows = self.svd_block_rows_forced
            if rows is None:
                rows = 16
                while rows > 1 and 2 * rows * (lmax + kmax) * 16 > 200 * 1024:
                    rows //= 2
            nb = max(2, 2 * ((kmax + 2 * rows - 1) // (2 * rows)))
            geo = self._jac_geo[key] = (rows, nb, 30 if nb == 2 else 1, lmax, kmax)
        return geo

    # Rows of the rotated matrix whose squared norm is below GRAM_TAU * s_max^2 are re-diagonalised at their own scale (second level).
    # Only the rows a 1e-12 discarded-weight decision can depend on need that: an unselected row contributes >= GRAM_TAU * s_max^2 to
    # every cumulative sum it is part of, and s_max^2 >= 1 / chi for a normalised state, so at 1e-8 those sums stay >= 7.8e-11 (chi = 128)
    # or 2e-11 (chi = 512) -- an order of magnitude above the threshold plus the first-level error bound (16 k eps s_max^2 <= 2e-12).
    # Round 1 used 1e-6, which re-diagonalised every singular value below 1e-3 s_max: a second eigen-solve on ~20 % of the rows of
    # EVERY gauge shift of the saturated C3 state for decisions that never came near the threshold.
    GRAM_TAU = float(os.environ.get("YB_GRAM_TAU", "1e-8"))

    def _svd_gram(self, mode: int, src: Buf, el: Ext, er: Ext, trunc, dist: str, mask, dims: tuple[int, int] | None = None) -> None:
        """M = U S Vh through the Hermitian eigen-problem of a Gram matrix.
        * split_two_site matrices (THETA): the Gram matrix is taken over the index that must become an isometry (rows for
          dist='right', columns for dist='left'), so the isometric factor is the eigenvector matrix itself and the other
          factor is the exact product E^H M -- no division by a singular value.
        * gauge shifts (ROWS / COLS): Gram matrix over the SMALL side (chi x chi instead of 2 chi x 2 chi).
        Gram eigenvalues carry an ABSOLUTE error ~ k eps s_max^2.  When the truncation threshold is small (the 1e-12 of the
        gauge shifts) the rows below GRAM_TAU * s_max^2 are re-diagonalised through their own Gram matrix (second level).
        Trajectories whose keep decision still lies within the error bound of the threshold are redone with the one-sided
        Jacobi (high relative accuracy)."""
        K, B, lay = self.K, self.B, self.lay
        sd = lay.svd_dim
        sq = sd * sd
        w, q = self._buf(self.svd_w, sq), self._buf(self.svd_q, sq)
        g, qh, zt = self._buf(self.heev_g, sq), self._buf(self.heev_q, sq), self._buf(self.heev_z, sq)
        tm, thr, mb, mk = trunc
        orient = 0 if mode != SVD_THETA else (SVD_ORIENT_ROWS if dist == "right" else SVD_ORIENT_COLS)
        K.svd_prepare(B, mask, mode | orient, el, er, self.d, src, w, q, self.svd_meta, self.svd_rot, self.counter, self.status_scratch)     # (clears the counter of svd_ambiguous and the scratch status)
        ek, eln = Ext(self.svd_meta, 0, SVD_META), Ext(self.svd_meta, 1, SVD_META)
        K.gemm(OP_GRAM, B, 1, mask, ek, eln, None, 1, 1, 1, w, w, g, 0, sd, sd)
        kdim = sd if dims is None else max(1, min(sd, min(dims) if mode != SVD_THETA else (dims[0] if dist == "right" else dims[1])))      # bound of k = rows of W
        # a split capped at max_bond < k can only keep the eigenvectors of the max_bond largest Gram eigenvalues: back-transform and
        # rotate only those rows; the singular values behind them (discarded weight) come from the eigenvalues (yb_heev_rows_top)
        top = mb if (self.heev_rows and self.heev_top and mode == SVD_THETA and 0 < mb < kdim and not (0.0 < thr < 1e-7)) else 0
        if top:
            K.heev_rows_top(B, mask, ek, g, q, self.heev_zr, sq, self.heev_d, self.heev_e, self.heev_tau, sd, self.svd_meta, self.status_scratch, kdim,
                            top, self.heev_order, self.heev_ds, self.heev_ntop)
            K.gemm(OP_ROTATE_TOP, B, 1, mask, ek, eln, Ext(self.heev_ntop, 0, 1), 1, 1, 1, q, w, g, 0, min(sd, top), sd)
            K.svd_finalize_top(B, mask, self.svd_meta, tm, thr, mb, mk, DIST[dist], g, q, self.svd_s, sd, Ext(self.keep, 0, 1),
                               self._buf(self.svd_u, sq), self._buf(self.svd_v, sq), self.status_scratch, sd, self.heev_ds, sd, top)
            self.stats["gram_svd_calls"] += 1
            self.stats["heev_top_calls"] += 1
            K.svd_ambiguous(B, mask, self.svd_meta, self.svd_s, sd, tm, thr, False, self.GRAM_TAU, 16.0 * sd * 2.220446049250313e-16,
                            self.status_scratch, self.amb, self.counter)
            nfb = K.sync_int(self.counter)
            if nfb > 0:
                self.stats["gram_fallbacks"] += nfb
                self._svd_jacobi(mode, src, el, er, trunc, dist, self.amb.clone(), dims)
            return
        if self.heev_rows:                               # blocked tridiagonalisation, eigenvector rows straight from the reflectors
            K.heev_rows(B, mask, ek, g, q, self.heev_zr, sq, self.heev_d, self.heev_e, self.heev_tau, sd, self.svd_meta, self.status_scratch, kdim)
        else:
            K.heev(B, mask, ek, g, qh, zt, self.heev_zr, sq, self.heev_d, self.heev_e, self.heev_tau, sd, self.svd_meta, self.status_scratch, sd)
            K.gemm(OP_EVEC, B, 1, mask, ek, None, None, 1, 1, 1, zt, qh, q, 0, sd, sd)
        K.gemm(OP_ROTATE, B, 1, mask, ek, eln, None, 1, 1, 1, q, w, g, 0, sd, sd)          # Wf -> heev_g
        # second level only where a threshold has to be resolved below the Gram accuracy; a truncation-free factorisation (thr == 0:
        # the QR-equivalent gauge moves) needs no singular value at all -- U is the unitary eigenvector matrix, S Vh the exact product
        two_level = 0.0 < thr < 1e-7
        if two_level:
            ws, qs, g2, hs = w, self._buf(self.svd_u, sq), self._buf(self.svd_v, sq), self._buf(self.heev_s, sq)
            K.svd_select(B, mask, self.svd_meta, g, self.GRAM_TAU, self.sel_n, self.sel_idx, sd)
            K.svd_gather(B, mask, self.svd_meta, self.sel_n, self.sel_idx, sd, 0, g, q, ws, qs)
            en = Ext(self.sel_n, 0, 1)
            K.gemm(OP_GRAM, B, 1, mask, en, eln, None, 1, 1, 1, ws, ws, g2, 0, sd, sd)
            if self.heev_rows:
                K.heev_rows(B, mask, en, g2, zt, self.heev_zr, sq, self.heev_d, self.heev_e, self.heev_tau, sd, None, self.status_scratch, kdim)
                e2 = zt
            else:
                K.heev(B, mask, en, g2, qh, zt, self.heev_zr, sq, self.heev_d, self.heev_e, self.heev_tau, sd, None, self.status_scratch, sd)
                K.gemm(OP_EVEC, B, 1, mask, en, None, None, 1, 1, 1, zt, qh, g2, 0, sd, sd)
                e2 = g2
            K.gemm(OP_ROTATE, B, 1, mask, en, eln, None, 1, 1, 1, e2, ws, hs, 0, sd, sd)
            K.gemm(OP_ROTATE, B, 1, mask, en, ek, None, 1, 1, 1, e2, qs, ws, 0, sd, sd)
            K.svd_gather(B, mask, self.svd_meta, self.sel_n, self.sel_idx, sd, 1, g, q, hs, ws)
        K.svd_finalize(B, mask, self.svd_meta, tm, thr, mb, mk, DIST[dist], g, q, self.svd_s, sd, Ext(self.keep, 0, 1),
                       self._buf(self.svd_u, sq), self._buf(self.svd_v, sq), self.status_scratch, sd)
        self.stats["gram_svd_calls"] += 1
        # ---- keep decisions within the error bound of the threshold -> Jacobi (one kernel + one host read per factorisation)
        K.svd_ambiguous(B, mask, self.svd_meta, self.svd_s, sd, tm, thr, two_level, self.GRAM_TAU, 16.0 * sd * 2.220446049250313e-16,
                        self.status_scratch, self.amb, self.counter)
        nfb = K.sync_int(self.counter)
        if nfb > 0:
            self.stats["gram_fallbacks"] += nfb
            self._svd_jacobi(mode, src, el, er, trunc, dist, self.amb.clone(), dims)

    def _svd_jacobi(self, mode: int, src: Buf, el: Ext, er: Ext, trunc, dist: str, mask, dims: tuple[int, int] | None = None) -> None:
        """prepare -> cyclic Jacobi sweeps until every trajectory is clean -> finalize (keep, U f(S), g(S) Vh)."""
        K, B, lay = self.K, self.B, self.lay
        sd = lay.svd_dim
        w, q = self._buf(self.svd_w, sd * sd), self._buf(self.svd_q, sd * sd)
        K.svd_prepare(B, mask, mode, el, er, self.d, src, w, q, self.svd_meta, self.svd_rot)
        rows, nb, inner, max_len, max_k = self._jacobi_geometry(dims)
        self.stats["svd_calls"] += 1
        for _sweep in range(60):
            for rnd in range(nb - 1):
                K.svd_round(B, mask, self.svd_meta, rows, nb, rnd, rnd == 0, inner, w, q, self.svd_rot, max_len, max_k, self.counter)      # (clears the counter of svd_check)
            K.svd_check(B, mask, self.svd_meta, self.svd_rot, self.counter)
            self.stats["svd_sweeps"] += 1
            if K.sync_int(self.counter) == 0:
                break
        tm, thr, mb, mk = trunc
        K.svd_finalize(B, mask, self.svd_meta, tm, thr, mb, mk, DIST[dist], w, q, self.svd_s, sd, Ext(self.keep, 0, 1),
                       self._buf(self.svd_u, sd * sd), self._buf(self.svd_v, sd * sd), self.status, sd)

    def split_theta(self, st: BatchedState, i: int, dist: str, trunc, mask=None, fast: bool = False) -> None:
        """split_two_site (decompositions.py:105-185): theta workspace -> sites i, i+1; bond i+1 <- keep."""
        lay, K, B, d = self.lay, self.K, self.B, self.d
        sd = lay.svd_dim
        kcap = (d * (lay.cap[i] if dist == "right" else lay.cap[i + 2])) if fast else 0
        self._svd(SVD_THETA, self._buf(self.theta, lay.theta_slot), st.ext(i), st.ext(i + 2), trunc, dist, mask, kcap, (d * lay.cap[i], d * lay.cap[i + 2]))
        kx = Ext(self.keep, 0, 1)
        K.copy_packed(B, mask, st.ext(i), kx, d, self._buf(self.svd_u, sd * sd), st.site(i), lay.slot[i])
        K.unpack_vh(B, mask, kx, st.ext(i + 2), d, self._buf(self.svd_v, sd * sd), st.site(i + 1), lay.slot[i + 1], st.ext(i + 1))      # bond i+1 <- keep

    def shift_right(self, st: BatchedState, i: int, trunc, mask=None, fast: bool = False) -> None:
        """Centre i -> i+1 (mps.py:719-769).  The reference merges the pair and splits it; because site i+1 is a right
        isometry the SVD of the centre tensor alone gives the same singular values / keep decision."""
        lay, K, B, d = self.lay, self.K, self.B, self.d
        sd = lay.svd_dim
        dims = (d * lay.cap[i], lay.cap[i + 1])
        kcap = min(d * lay.cap[i], lay.cap[i + 1]) if fast else 0

        def install(m):                                      # U -> site i, (S Vh) absorbed into site i+1, bond i+1 <- keep
            kx = Ext(self.keep, 0, 1)
            K.copy_packed(B, m, st.ext(i), kx, d, self._buf(self.svd_u, sd * sd), st.site(i), lay.slot[i])
            tmp = self._buf(self.tmp, lay.nloc_max)
            K.gemm(OP_ABSORB_R, B, d, m, kx, st.ext(i + 1), st.ext(i + 2), d, 1, 1, self._buf(self.svd_v, sd * sd), st.site(i + 1), tmp, 0,
                   lay.cap[i + 1], lay.cap[i + 2])
            K.copy_packed(B, m, kx, st.ext(i + 2), d, tmp, st.site(i + 1), lay.slot[i + 1], st.ext(i + 1))

        if self._cholqr_applies(trunc, "right", dims, kcap):
            self._shift_cholqr(("shr", id(st), i), SVD_ROWS, st.site(i), st.ext(i), st.ext(i + 1), trunc, "right", mask, dims, install)
            return
        self._svd(SVD_ROWS, st.site(i), st.ext(i), st.ext(i + 1), trunc, "right", mask, kcap, dims)
        install(mask)

    def shift_left(self, st: BatchedState, i: int, trunc, mask=None, fast: bool = False) -> None:
        """Centre i -> i-1 (mps.py:771-788), reduced form (site i-1 must be a left isometry)."""
        lay, K, B, d = self.lay, self.K, self.B, self.d
        sd = lay.svd_dim
        dims = (lay.cap[i], d * lay.cap[i + 1])
        kcap = min(lay.cap[i], d * lay.cap[i + 1]) if fast else 0

        def install(m):                                      # (U S) absorbed into site i-1, Vh -> site i, bond i <- keep
            kx = Ext(self.keep, 0, 1)
            tmp = self._buf(self.tmp, lay.nloc_max)
            K.gemm(OP_ABSORB_L, B, d, m, st.ext(i - 1), st.ext(i), kx, d, 1, 1, st.site(i - 1), self._buf(self.svd_u, sd * sd), tmp, 0,
                   lay.cap[i - 1], lay.cap[i])
            K.copy_packed(B, m, st.ext(i - 1), kx, d, tmp, st.site(i - 1), lay.slot[i - 1])
            K.unpack_vh(B, m, kx, st.ext(i + 1), d, self._buf(self.svd_v, sd * sd), st.site(i), lay.slot[i], st.ext(i))

        if self._cholqr_applies(trunc, "left", dims, kcap):
            self._shift_cholqr(("shl", id(st), i), SVD_COLS, st.site(i), st.ext(i), st.ext(i + 1), trunc, "left", mask, dims, install)
            return
        self._svd(SVD_COLS, st.site(i), st.ext(i), st.ext(i + 1), trunc, "left", mask, kcap, dims)
        install(mask)

    def _shift_cholqr(self, key, mode: int, src: Buf, el: Ext, er: Ext, trunc, dist: str, mask, dims, install) -> None:
        """One keep-all gauge shift on the Cholesky-QR route as a fixed launch sequence (CUDA graph from its third use on): the
        factorisation, the mask of the accepted trajectories, the launches that install their factors.  One host read afterwards;
        flagged trajectories (their site tensor is untouched) take the eigen-solver route and are installed under their own mask."""
        K, B = self.K, self.B
        gm = None
        if mask is not None:                                 # a captured sequence needs the mask at a fixed address
            self.g_mask.copy_(mask)
            gm = self.g_mask

        def seq():
            self._cholqr_chain(mode, src, el, er, trunc, dist, gm, dims)
            K.mask_andnot(B, gm, self.amb, self.g_work)
            install(self.g_work)

        self._run_graphed(key + (float(trunc[1]), mask is not None), seq)
        self.stats["cholqr_calls"] += 1
        nfb = K.sync_int(self.counter)
        if nfb > 0:
            self.stats["cholqr_fallbacks"] += nfb
            redo = self.amb.clone()
            self._svd_gram(mode, src, el, er, trunc, dist, redo, dims)
            install(redo)

    def normalize_site(self, st: BatchedState, i: int, mask=None) -> None:
        """QR at a chain end with R dropped == divide the (1-column) centre tensor by its norm (mps.py:736-746)."""
        K, B = self.K, self.B
        K.norm2(B, mask, st.ext(i), st.ext(i + 1), self.d, st.site(i), self.norm_buf)
        K.scale(B, mask, st.ext(i), st.ext(i + 1), self.d, self.norm_buf, st.site(i), self.lay.slot[i], True)

    # ------------------------------------------------------------------------------------------------------------------
    # TDVP sweeps (integrators.py:44-291)
    # ------------------------------------------------------------------------------------------------------------------
    def _substep_dt(self, scale: float) -> float:
        return scale if self.p.digital else float(self.p.dt) * scale

    def tdvp(self, st: BatchedState) -> None:
        """tdvp (tdvp.py:69-112)."""
        self.K.phase = "tdvp"
        mode = self.p.tdvp_mode
        if mode in ("2site", "dynamic") and self.L == 1:
            mode = "1site"
        plan = [1.0 / self.p.tdvp_sweeps] * self.p.tdvp_sweeps
        if mode == "2site":
            self.sweep_2site(st, plan)
        elif mode == "1site":
            for s in plan:
                self.sweep_1site(st, s)
        elif mode == "dynamic":
            for s in plan:
                self.sweep_dynamic(st, s)
        else:
            raise ValueError(f'tdvp_mode must be one of ("1site", "2site", "dynamic"), got {mode!r}.')

    def sweep_2site(self, st: BatchedState, plan: Sequence[float], lo: int = 0, hi: int | None = None) -> None:
        """integrators.py:161-291 on sites lo..hi (default: the whole chain; a sub-range is a digital gate window whose
        boundary bonds keep their dimension and get identity environments, digital_tjm.py:360-405)."""
        lay, p = self.lay, self.p
        hi = self.L - 1 if hi is None else hi
        n = hi + 1
        trunc = self._trunc_args(p.trunc_mode, p.svd_threshold, p.max_bond_dim, p.tdvp_min_keep)
        th = self._buf(self.theta, lay.theta_slot)
        self.init_right_envs(st, lo, hi)
        for scale in plan:
            dt = self._substep_dt(scale)
            for i in range(lo, n - 2):
                self.merge(st, i)
                self.update_site(st, i, 2, th, 0.5 * dt)
                self.split_theta(st, i, "right", trunc, fast=True)
                self.env_update_left(st, i)
                self.update_site(st, i + 1, 1, st.site(i + 1), -0.5 * dt)
            i = n - 2
            self.merge(st, i)
            self.update_site(st, i, 2, th, dt)
            self.split_theta(st, i, "left", trunc, fast=True)
            self.env_update_right(st, i + 1)
            for i in reversed(range(lo, n - 2)):
                self.update_site(st, i + 1, 1, st.site(i + 1), -0.5 * dt)
                self.merge(st, i)
                self.update_site(st, i, 2, th, 0.5 * dt)
                self.split_theta(st, i, "left", trunc, fast=True)
                self.env_update_right(st, i + 1)

    def sweep_1site(self, st: BatchedState, scale: float) -> None:
        n, lay, K, B, d = self.L, self.lay, self.K, self.B, self.d
        keep_all = self._trunc_args("discarded_weight", 0.0, None, 1)      # QR: no truncation (integrators.py:100,131)
        dt = self._substep_dt(scale)
        sd = lay.svd_dim
        kx = Ext(self.keep, 0, 1)
        cbuf = self._buf(self.theta, lay.theta_slot)
        tmp = self._buf(self.tmp, lay.nloc_max)
        self.init_right_envs(st)
        if n == 1:
            self.update_site(st, 0, 1, st.site(0), dt)
            return
        for i in range(n - 1):
            self.update_site(st, i, 1, st.site(i), 0.5 * dt)
            self._svd(SVD_ROWS, st.site(i), st.ext(i), st.ext(i + 1), keep_all, "right", None, 0, (d * lay.cap[i], lay.cap[i + 1]))
            K.copy_packed(B, None, st.ext(i), kx, d, self._buf(self.svd_u, sd * sd), st.site(i), lay.slot[i])
            K.copy_packed(B, None, kx, st.ext(i + 1), 1, self._buf(self.svd_v, sd * sd), cbuf, lay.theta_slot)      # C (k x r_old)
            r_old = st.bond[:, i + 1].clone()
            st.bond[:, i + 1] = self.keep                   # site i now has right bond k
            self.env_update_left(st, i)                     # lb[i+1] with extents (k, D, k)
            st.bond[:, i + 1] = r_old                       # rb[i] still has the old extent
            kt = self.keep.clone()
            ek, eo = Ext(kt, 0, 1), st.ext(i + 1)
            self.krylov(cbuf, cbuf, ek, eo, 1, lambda s_, d_, m, ek=ek, eo=eo, j=i + 1: self.bond_matvec(ek, eo, j, s_, d_, m, lay.cap[j], lay.cap[j]),
                        -0.5 * dt, None)
            K.gemm(OP_ABSORB_R, B, d, None, ek, eo, st.ext(i + 2), d, 1, 1, cbuf, st.site(i + 1), tmp, 0, lay.cap[i + 1], lay.cap[i + 2])
            K.copy_packed(B, None, ek, st.ext(i + 2), d, tmp, st.site(i + 1), lay.slot[i + 1])
            st.bond[:, i + 1] = kt
        self.update_site(st, n - 1, 1, st.site(n - 1), dt)
        for i in reversed(range(1, n)):
            self._svd(SVD_COLS, st.site(i), st.ext(i), st.ext(i + 1), keep_all, "left", None, 0, (lay.cap[i], d * lay.cap[i + 1]))
            K.unpack_vh(B, None, kx, st.ext(i + 1), d, self._buf(self.svd_v, sd * sd), st.site(i), lay.slot[i])
            K.copy_packed(B, None, st.ext(i), kx, 1, self._buf(self.svd_u, sd * sd), cbuf, lay.theta_slot)          # C (l_old x k)
            l_old = st.bond[:, i].clone()
            kt = self.keep.clone()
            st.bond[:, i] = kt
            self.env_update_right(st, i)                    # rb[i-1] with extents (k, D, k)
            st.bond[:, i] = l_old
            eo, ek = st.ext(i), Ext(kt, 0, 1)
            self.krylov(cbuf, cbuf, eo, ek, 1, lambda s_, d_, m, ek=ek, eo=eo, j=i: self.bond_matvec(eo, ek, j, s_, d_, m, lay.cap[j], lay.cap[j]),
                        -0.5 * dt, None)
            K.gemm(OP_ABSORB_L, B, d, None, st.ext(i - 1), eo, ek, d, 1, 1, st.site(i - 1), cbuf, tmp, 0, lay.cap[i - 1], lay.cap[i])
            K.copy_packed(B, None, st.ext(i - 1), ek, d, tmp, st.site(i - 1), lay.slot[i - 1])
            st.bond[:, i] = kt
            self.update_site(st, i - 1, 1, st.site(i - 1), 0.5 * dt)

    def _cap_bonds(self, st: BatchedState) -> None:
        """_cap_bonds / _sync_bond_dim (sweep_utils.py:110-164,274-294), analog path: every bond above the cap is merged with its
        pair and re-split ("sqrt", the user's truncation policy, cap = max_bond_dim, min_keep 1), in ascending bond order, for the
        trajectories that need it."""
        cap, n, p = self.p.max_bond_dim, self.L, self.p
        if cap is None or n < 2:
            return
        over = (st.bond[:, 1:n] > cap).any(dim=0).cpu().tolist()
        trunc = self._trunc_args(p.trunc_mode, p.svd_threshold, cap, 1)
        for c, flag in enumerate(over):
            if flag:
                j = c + 1
                mask = (st.bond[:, j] > cap).to(torch.int32)
                self.merge(st, j - 1, mask)
                self.split_theta(st, j - 1, "sqrt", trunc, mask)

    def sweep_dynamic(self, st: BatchedState, scale: float) -> None:
        """sweep_dynamic (integrators.py:294-511) on the analog path (no fixed-chi renormalisation, sweep_utils.py:191-202): per
        site every trajectory takes the 1-site update with a QR hand-over when ITS bond has reached the cap, else the 2-site update
        with an uncapped split -- two complementary per-trajectory masks, the environment update in between is common to both.
        The QR hand-overs are the truncation-free SVD surrogate of ``sweep_1site`` (a gauge of the same factorisation).

        Parity: exact against the reference as long as no trajectory takes the 1-site branch on the right-to-left pass (uncapped
        runs, or bonds that stay below the cap) -- fixture heis6_dynamic_nocap_order2.  On that branch the reference transposes the
        bond matrix twice (``left_qr`` already returns it as (left, new), decompositions.py:82; integrators.py:458 transposes again),
        so the Householder factor R itself is evolved and absorbed where R^T belongs, and above the cap R[:, :cap] is kept
        (:450-453).  Shapes agree because k = l there, but the result is not a gauge-invariant function of the state: it depends on
        the phases LAPACK's SVD / QR happened to choose earlier in the sweep, and it is not a TDVP step any more (L = 2, chi = 2,
        dt = 0.1: overlap 0.962 with exp(-i dt H)|psi> where 1-site TDVP is exact).  The oracle restates it literally and is pinned
        to the reference; the engine evolves R^T -- the projector-splitting step the code was written to take -- and truncates to
        the cap in Schmidt space, so in the saturated regime it agrees with exact evolution, not with the reference's numbers."""
        n, lay, K, B, d, p = self.L, self.lay, self.K, self.B, self.d, self.p
        cap = p.max_bond_dim
        keep_all = self._trunc_args("discarded_weight", 0.0, None, 1)
        cap_only = self._trunc_args("discarded_weight", 0.0, cap, 1)
        dt = self._substep_dt(scale)
        sd = lay.svd_dim
        kx = Ext(self.keep, 0, 1)
        cbuf = self._buf(self.theta, lay.theta_slot)
        tmp = self._buf(self.tmp, lay.nloc_max)
        self._cap_bonds(st)
        self.init_right_envs(st)

        def dyn(j: int):                                    # uncapped split; the slot capacity is the only bound
            return self._trunc_args(p.trunc_mode, p.svd_threshold, lay.cap[j], p.tdvp_min_keep)

        def branches(col: torch.Tensor):
            m1 = (col >= cap) if cap is not None else torch.zeros_like(col, dtype=torch.bool)
            any1, any2 = torch.stack([m1.any(), (~m1).any()]).cpu().tolist()
            return m1, m1.to(torch.int32), (~m1).to(torch.int32), bool(any1), bool(any2)

        # ----- left to right (:341-427) -----
        for i in range(n):
            m1, m1i, m2i, any1, any2 = branches(st.bond[:, i + 1])
            if any1:
                self.update_site(st, i, 1, st.site(i), 0.5 * dt, m1i)
            if i == n - 1:
                continue
            k1 = None
            if any1:
                self._svd(SVD_ROWS, st.site(i), st.ext(i), st.ext(i + 1), keep_all, "right", m1i, 0, (d * lay.cap[i], lay.cap[i + 1]))
                K.copy_packed(B, m1i, st.ext(i), kx, d, self._buf(self.svd_u, sd * sd), st.site(i), lay.slot[i])
                K.copy_packed(B, m1i, kx, st.ext(i + 1), 1, self._buf(self.svd_v, sd * sd), cbuf, lay.theta_slot)      # C (k x r_old)
                k1 = self.keep.clone()
            if any2:
                th = self.merge(st, i, m2i)
                self.update_site(st, i, 2, th, 0.5 * dt, m2i)
                self.split_theta(st, i, "right", dyn(i + 1), m2i)
            r_old = st.bond[:, i + 1].clone()
            if any1:
                st.bond[:, i + 1] = torch.where(m1, k1, r_old)
            self.env_update_left(st, i)
            if any1:
                st.bond[:, i + 1] = r_old
                kt = torch.where(m1, k1, r_old)
                ek, eo = Ext(kt, 0, 1), st.ext(i + 1)
                self.krylov(cbuf, cbuf, ek, eo, 1, lambda s_, d_, m, ek=ek, eo=eo, j=i + 1: self.bond_matvec(ek, eo, j, s_, d_, m, lay.cap[j], lay.cap[j]),
                            -0.5 * dt, m1i)
                K.gemm(OP_ABSORB_R, B, d, m1i, ek, eo, st.ext(i + 2), d, 1, 1, cbuf, st.site(i + 1), tmp, 0, lay.cap[i + 1], lay.cap[i + 2])
                K.copy_packed(B, m1i, ek, st.ext(i + 2), d, tmp, st.site(i + 1), lay.slot[i + 1])
                st.bond[:, i + 1] = kt
            if any2:
                if i == n - 2:
                    self.env_update_right(st, n - 1)
                else:
                    self.update_site(st, i + 1, 1, st.site(i + 1), -0.5 * dt, m2i)
        # ----- right to left (:430-506) -----
        for i in reversed(range(n)):
            m1, m1i, m2i, any1, any2 = branches(st.bond[:, i])
            if any1:
                self.update_site(st, i, 1, st.site(i), 0.5 * dt, m1i)
            if i == 0:
                continue
            k1 = None
            if any1:
                # QR hand-over to the left by the truncation-free SVD surrogate of sweep_1site; where the new bond would exceed the cap
                # (the left bond came out of an uncapped split) it is truncated to the cap in Schmidt space.  NOT what the reference
                # computes here -- see the docstring.
                self._svd(SVD_COLS, st.site(i), st.ext(i), st.ext(i + 1), cap_only, "left", m1i, 0, (lay.cap[i], d * lay.cap[i + 1]))
                K.unpack_vh(B, m1i, kx, st.ext(i + 1), d, self._buf(self.svd_v, sd * sd), st.site(i), lay.slot[i])
                K.copy_packed(B, m1i, st.ext(i), kx, 1, self._buf(self.svd_u, sd * sd), cbuf, lay.theta_slot)          # C (l_old x k)
                k1 = self.keep.clone()
            if any2:
                th = self.merge(st, i - 1, m2i)
                self.update_site(st, i - 1, 2, th, 0.5 * dt, m2i)
                self.split_theta(st, i - 1, "left", dyn(i), m2i)
            l_old = st.bond[:, i].clone()
            if any1:
                st.bond[:, i] = torch.where(m1, k1, l_old)
            self.env_update_right(st, i)
            if any1:
                st.bond[:, i] = l_old
                kt = torch.where(m1, k1, l_old)
                eo, ek = st.ext(i), Ext(kt, 0, 1)
                self.krylov(cbuf, cbuf, eo, ek, 1, lambda s_, d_, m, ek=ek, eo=eo, j=i: self.bond_matvec(eo, ek, j, s_, d_, m, lay.cap[j], lay.cap[j]),
                            -0.5 * dt, m1i)
                K.gemm(OP_ABSORB_L, B, d, m1i, st.ext(i - 1), eo, ek, d, 1, 1, st.site(i - 1), cbuf, tmp, 0, lay.cap[i - 1], lay.cap[i])
                K.copy_packed(B, m1i, st.ext(i - 1), ek, d, tmp, st.site(i - 1), lay.slot[i - 1])
                st.bond[:, i] = kt
            if any2 and i != 1:
                self.update_site(st, i - 1, 1, st.site(i - 1), -0.5 * dt, m2i)

    # ------------------------------------------------------------------------------------------------------------------
    # dissipation (dissipation.py:50-183)
    # ------------------------------------------------------------------------------------------------------------------
    def _dissipators(self, dt: float):
        """Per-site local factors for time step dt: (op2x2 | None, pair4x4 | None).  Identical for every trajectory, so
        scipy.linalg.expm of the 2x2 / 4x4 generators runs once per (dt) on the host."""
        if dt in self._diss_cache:
            return self._diss_cache[dt]
        from .noise import is_pauli
        n = self.L
        out = []
        for i in range(n):
            op1 = np.eye(2, dtype=np.complex128)
            touched = False
            gen = np.zeros((2, 2), dtype=np.complex128)
            has_one, all_pauli = False, True
            for pr in self.noise:
                if len(pr["sites"]) == 1 and pr["sites"][0] == i:
                    has_one = True
                    if is_pauli(pr):
                        gen += pr["strength"] * np.eye(2)
                    else:
                        all_pauli = False
                        m = np.asarray(pr["matrix"], dtype=np.complex128)
                        gen += pr["strength"] * (m.conj().T @ m)
            if has_one:
                touched = True
                op1 = (np.exp(-0.5 * dt * float(gen[0, 0].real)) * np.eye(2)) if all_pauli else scipy.linalg.expm(-0.5 * dt * gen)
            pair = None
            scalar2 = 1.0
            if i != 0:
                gen2 = np.zeros((4, 4), dtype=np.complex128)
                has_adj, all_pauli2 = False, True
                for pr in self.noise:
                    if len(pr["sites"]) == 2 and pr["sites"][1] == i:
                        if abs(pr["sites"][1] - pr["sites"][0]) > 1:
                            if not is_pauli(pr):
                                raise NotImplementedError("Non-Pauli Long-range processes are not implemented yet")
                            scalar2 *= np.exp(-0.5 * dt * pr["strength"])
                        else:
                            has_adj = True
                            if is_pauli(pr):
                                gen2 += pr["strength"] * np.eye(4)
                            else:
                                all_pauli2 = False
                                m = np.asarray(pr["matrix"], dtype=np.complex128)
                                gen2 += pr["strength"] * (m.conj().T @ m)
                if has_adj:
                    if all_pauli2:
                        scalar2 *= np.exp(-0.5 * dt * float(gen2[0, 0].real))
                    else:
                        pair = scipy.linalg.expm(-0.5 * dt * gen2)
            out.append((op1 if touched else None, None if scalar2 == 1.0 else scalar2, pair))
        self._diss_cache[dt] = out
        return out

    def apply_dissipation(self, st: BatchedState, dt: float, start: int = 0) -> None:
        """dissipation.py:50-183; ``start`` = current orthogonality centre (0 on the analog path, the right site of the
        gate pair after a TEBD gate)."""
        lay, K, B, d, n, p = self.lay, self.K, self.B, self.d, self.L, self.p
        K.phase = "dissipation"
        if not self.noise_active:
            keep_all = self._trunc_args("discarded_weight", 0.0, None, 1)
            for i in reversed(range(1, start + 1)):         # dissipation.py:79-86: centre -> 0 with QR shifts
                self.shift_left(st, i, keep_all, fast=True)
            self.normalize_site(st, 0)                      # QR at site 0, R dropped
            return
        shift = self._trunc_args("discarded_weight", SHIFT_THRESHOLD, None, 1)
        user = self._trunc_args(p.trunc_mode, p.svd_threshold, p.max_bond_dim, 1)
        for i in range(start, n - 1):                       # centre -> L-1 with SVD shifts (:88-92)
            self.shift_right(st, i, shift, fast=True)
        facs = self._dissipators(dt)
        for i in reversed(range(n)):
            op1, scalar2, pair = facs[i]
            if op1 is not None:
                op_t = self._const(op1)
                K.apply_local_op(B, None, None, 0, st.ext(i), st.ext(i + 1), d, op_t, 0, st.site(i), lay.cap[i] * lay.cap[i + 1])
            if scalar2 is not None:
                op_t = self._const(scalar2 * np.eye(2, dtype=np.complex128))
                K.apply_local_op(B, None, None, 0, st.ext(i), st.ext(i + 1), d, op_t, 0, st.site(i), lay.cap[i] * lay.cap[i + 1])
            if pair is not None:
                th = self.merge(st, i - 1)
                op_t = self._const(pair)
                K.apply_local_op(B, None, None, 0, st.ext(i - 1), st.ext(i + 1), d * d, op_t, 0, th, lay.cap[i - 1] * lay.cap[i + 1])
                self.split_theta(st, i - 1, "right", user, fast=True)
            if i != 0:
                self.shift_left(st, i, shift, fast=True)

    # ------------------------------------------------------------------------------------------------------------------
    # local reduced density matrices by a left-to-right transfer sweep (observables + jump weights)
    # ------------------------------------------------------------------------------------------------------------------
    def rho_sweep(self, st: BatchedState, need1: Sequence[int], need2: Sequence[int], mask=None):
        """g1[i, b, o, p] = <psi| |o><p|_i |psi> for i in need1 and the two-site analogue g2[i] for pairs (i, i+1), i in need2,
        valid when every site right of the queried one is a right isometry (centre 0).  Replaces the reference's
        copy + QR-shift + local contraction (mps.py:1178-1234, stochastic_process.py:139-176) by two GEMMs per site.  The local
        density matrices of all sites land in two device tables ([L, B, d, d] and [L - 1, B, d^2, d^2]; rows outside need1 /
        need2 are stale), so one batched contraction per observable kind follows instead of one per observable."""
        lay, K, B, d, n = self.lay, self.K, self.B, self.d, self.L
        need1, need2 = set(need1), set(need2)
        last = max([*need1, *[i + 1 for i in need2], 0])
        if self._rdm1_all is None:
            self._rdm1_all = torch.zeros((n, B, d * d), dtype=C128, device=self.dev)
        if need2 and self._rdm2_all is None:
            self._rdm2_all = torch.zeros((max(n - 1, 1), B, d ** 4), dtype=C128, device=self.dev)
        x = self._buf(self.T1, lay.t_slot)
        K.env_identity(B, mask, st.ext(0), 1, Buf(self.lenv, lay.env_off[0], lay.env_slot[0]), 1)
        for i in range(last + 1):
            e_i = Buf(self.lenv, lay.env_off[i], lay.env_slot[i])
            if i in need2:
                th = self.merge(st, i, mask)
                K.gemm(OP_LCONTRACT, B, d * d, mask, st.ext(i), st.ext(i + 2), None, d * d, 1, 1, e_i, th, x, 0, lay.cap[i], lay.cap[i + 2])
                K.local_rdm(B, mask, st.ext(i), st.ext(i + 2), d * d, th, x, self._rdm2_all[i])
            a = st.site(i)
            K.gemm(OP_LCONTRACT, B, d, mask, st.ext(i), st.ext(i + 1), None, d, 1, 1, e_i, a, x, 0, lay.cap[i], lay.cap[i + 1])
            if i in need1:
                K.local_rdm(B, mask, st.ext(i), st.ext(i + 1), d, a, x, self._rdm1_all[i])
            if i < last:
                e_n = Buf(self.lenv, lay.env_off[i + 1], lay.env_slot[i + 1])
                K.gemm(OP_ENVL3, B, 1, mask, st.ext(i), st.ext(i + 1), None, d, 1, 1, x, a, e_n, 1, lay.cap[i + 1], lay.cap[i + 1])
        g2 = self._rdm2_all.view(-1, B, d * d, d * d) if self._rdm2_all is not None else None
        return self._rdm1_all.view(n, B, d, d), g2

    def _observable_tables(self) -> dict:
        """Device tables of the local observables, built once: operator stacks and site / column indices per kind."""
        if self._obs_tab is None:
            one = [(k, m, s[0]) for k, (m, s) in enumerate(self.observables) if not isinstance(m, str) and len(s) == 1]
            two = [(k, m, s[0]) for k, (m, s) in enumerate(self.observables) if not isinstance(m, str) and len(s) == 2]
            for _, (m, s) in enumerate(self.observables):
                if not isinstance(m, str) and len(s) == 2 and s[1] != s[0] + 1:
                    raise AssertionError("Only nearest-neighbor observables are currently implemented.")
            tab = {"need1": sorted({s for _, _, s in one}), "need2": sorted({s for _, _, s in two}), "any": bool(one or two)}
            items = [(k, m, s, 0) for k, m, s in one] + [(k, m, s, 1) for k, m, s in two]
            if items:                                       # tables of yb_observables: result column, site, one- / two-site, operator (16 slots)
                mats = np.zeros((len(items), 16), dtype=np.complex128)
                for row, (_, m, _, tw) in enumerate(items):
                    mats[row, : 16 if tw else 4] = np.asarray(m, dtype=np.complex128).reshape(-1)
                tab["n"] = len(items)
                tab["col"] = torch.tensor([k for k, _, _, _ in items], dtype=torch.int32, device=self.dev)
                tab["site"] = torch.tensor([s for _, _, s, _ in items], dtype=torch.int32, device=self.dev)
                tab["two"] = torch.tensor([tw for _, _, _, tw in items], dtype=torch.int32, device=self.dev)
                tab["mats"] = torch.from_numpy(mats).to(self.dev)
            self._obs_tab = tab
        return self._obs_tab

    def evaluate_observables(self, st: BatchedState) -> torch.Tensor:
        """[B, n_obs] expectation values in the given (sorted) observable order (mps.py:1178-1234).  A value whose imaginary
        part violates the reference's ``assert exp.imag < 1e-13`` (mps.py:1233) raises status bit 8 for that trajectory."""
        tab = self._observable_tables()
        self.K.phase = "observables"
        res = torch.zeros((self.B, len(self.observables)), dtype=torch.float64, device=self.dev)
        if tab["any"]:
            self.rho_sweep(st, tab["need1"], tab["need2"])
            self.K.observables(self.B, tab["n"], tab["col"], tab["site"], tab["two"], tab["mats"], self._rdm1_all,
                               self._rdm2_all if tab["need2"] else None, res, res.shape[1], self.status)
        self._spectra_now = {}
        for k, (mat, sites) in enumerate(self.observables):
            if not isinstance(mat, str):
                continue
            if mat == "entropy":
                res[:, k] = self.entropy_of_cut(st, sites[0])
            elif mat == "schmidt_spectrum":              # vector-valued: the row holds NaN, the spectrum goes to self.spectra[(k, column)]
                self._spectra_now[k] = self.schmidt_of_cut(st, sites[0])
                res[:, k] = float("nan")
            else:
                res[:, k] = self.project_onto_bitstring(st, mat[4:])
        return res

    SCHMIDT_PAD = 500                                    # mps.py:660: the spectrum is padded / truncated to 500 entries

    def schmidt_of_cut(self, st: BatchedState, i: int) -> torch.Tensor:
        """MPS.get_schmidt_spectrum (mps.py:644-678): singular values of A_i A_{i+1} read as (d l) x (d r) on the tensors as stored,
        NaN-padded to 500 entries; a bond of dimension 1 gives [1, NaN, ...] (:664-667).  [B, 500] float64.  The one-sided Jacobi
        (high relative accuracy for every singular value) is used whatever the size -- this is a diagnostic, not the hot path."""
        lay, B = self.lay, self.B
        sd = lay.svd_dim
        self.merge(st, i)
        keep_all = self._trunc_args("discarded_weight", 0.0, None, 1)
        self._svd(SVD_THETA, self._buf(self.theta, lay.theta_slot), st.ext(i), st.ext(i + 2), keep_all, "right", None, 0,
                  (self.d * lay.cap[i], self.d * lay.cap[i + 2]))
        meta = self.svd_meta.view(B, SVD_META)
        nsv = torch.minimum(meta[:, 3], meta[:, 4])
        n = min(self.SCHMIDT_PAD, sd)
        out = torch.full((B, self.SCHMIDT_PAD), float("nan"), dtype=torch.float64, device=self.dev)
        sv = self.svd_s.view(B, sd)[:, :n]
        out[:, :n] = torch.where(torch.arange(n, device=self.dev)[None, :] < nsv[:, None], sv, torch.full_like(sv, float("nan")))
        trivial = st.bond[:, i + 1] == 1
        if bool(trivial.any()):
            out[trivial] = float("nan")
            out[trivial, 0] = 1.0
        return out

    def entropy_of_cut(self, st: BatchedState, i: int) -> torch.Tensor:
        """MPS.get_entropy (mps.py:604-639): von Neumann entropy of the singular values of A_i A_{i+1} read as (d l) x (d r), on
        the tensors AS STORED (the reference does not move the gauge centre for this meta-observable).  [B] float64."""
        lay, B = self.lay, self.B
        sd = lay.svd_dim
        self.merge(st, i)
        keep_all = self._trunc_args("discarded_weight", 0.0, None, 1)
        self._svd(SVD_THETA, self._buf(self.theta, lay.theta_slot), st.ext(i), st.ext(i + 2), keep_all, "right", None, self.d * lay.cap[i],
                  (self.d * lay.cap[i], self.d * lay.cap[i + 2]))
        meta = self.svd_meta.view(B, SVD_META)
        nsv = torch.minimum(meta[:, 3], meta[:, 4])
        sv = self.svd_s.view(B, sd)
        s2 = torch.where(torch.arange(sd, device=self.dev)[None, :] < nsv[:, None], sv * sv, torch.zeros_like(sv))
        norm = s2.sum(dim=1)
        pr = s2 / torch.where(norm > 0, norm, torch.ones_like(norm))[:, None]
        ent = -(pr * torch.log(pr + float(np.finfo(np.float64).tiny))).sum(dim=1)
        return torch.where(norm > 0, ent, torch.zeros_like(ent))

    def project_onto_bitstring(self, st: BatchedState, bits: str) -> torch.Tensor:
        """MPS.project_onto_bitstring (mps.py:1495-1537): probability of a computational-basis outcome by sequential projection.
        The carry (1 x chi row vector) lives in the theta workspace; per site one ABSORB_R contraction, a slice selection, a norm
        and a scaling -- the state itself is not touched (the reference works on a deep copy).  [B] float64."""
        K, B, d, lay = self.K, self.B, self.d, self.lay
        assert len(bits) == self.L, "Bitstring length must match number of sites"
        one = Ext(self._ones_i32, 0, 1)
        carry = self._buf(self.theta, lay.theta_slot)
        tmp = self._buf(self.tmp, lay.nloc_max)
        total = torch.ones(B, dtype=torch.float64, device=self.dev)
        for site, ch in enumerate(bits):
            bit = int(ch)
            assert 0 <= bit < d, f"Invalid state index {bit} at site {site}"
            er = st.ext(site + 1)
            if site == 0:
                K.copy_packed(B, None, st.ext(0), er, d, st.site(0), tmp, lay.slot[0])
            else:
                K.gemm(OP_ABSORB_R, B, d, None, one, st.ext(site), er, d, 1, 1, carry, st.site(site), tmp, 0, 1, lay.cap[site + 1])
            sel = np.zeros((d, d), dtype=np.complex128)
            sel[0, bit] = 1.0                                            # slice 0 <- slice `bit`, the rest <- 0
            K.apply_local_op(B, None, None, 0, one, er, d, self._const(sel), 0, tmp, lay.cap[site + 1])
            K.norm2(B, None, one, er, d, tmp, self.norm_buf)
            nrm = torch.sqrt(self.norm_buf)
            total = total * nrm
            self.scale_buf.copy_(torch.where(nrm > 0, 1.0 / torch.where(nrm > 0, nrm, torch.ones_like(nrm)), torch.zeros_like(nrm)))
            K.scale(B, None, one, er, d, self.scale_buf, tmp, d * lay.cap[site + 1])
            K.copy_packed(B, None, one, er, 1, tmp, carry, lay.cap[site + 1])
        return total * total

    def measure_shots(self, st: BatchedState, u: torch.Tensor) -> torch.Tensor:
        """MPS.measure_single_shot (mps.py:1282-1340) in the Z basis, u.shape[1] shots of every trajectory: sequential projective
        sampling from site 0 to L-1 with the centre at 0.  Per site the outcome probabilities are the diagonal of the local
        reduced density matrix of carry x A_site (``k_local_rdm``), the outcome follows numpy's ``Generator.choice``
        (``searchsorted(cdf, u, 'right')`` on the renormalised CDF), the selected slice, scaled by 1/sqrt(p), is carried into the
        next site with one ABSORB_R contraction.  u: [B, nshots, L] uniforms.  Returns [B, nshots] int64, bit i = outcome of site
        i (``sum(c << i)``, mps.py:1340).  The state is not touched (the reference samples a deep copy)."""
        K, B, d, lay, n = self.K, self.B, self.d, self.lay, self.L
        assert n <= 62, "outcomes are packed into int64 bit fields"
        nshots = int(u.shape[1])
        out = torch.zeros((B, nshots), dtype=torch.int64, device=self.dev)
        one = Ext(self._ones_i32, 0, 1)
        carry = self._buf(self.theta, lay.theta_slot)
        tmp = self._buf(self.tmp, lay.nloc_max)
        rows = torch.arange(B, device=self.dev)
        sel = torch.zeros((B, d * d), dtype=torch.complex128, device=self.dev)
        for shot in range(nshots):
            for site in range(n):
                er = st.ext(site + 1)
                if site == 0:
                    K.copy_packed(B, None, st.ext(0), er, d, st.site(0), tmp, lay.slot[0])
                else:
                    K.gemm(OP_ABSORB_R, B, d, None, one, st.ext(site), er, d, 1, 1, carry, st.site(site), tmp, 0, 1, lay.cap[site + 1])
                K.local_rdm(B, None, one, er, d, tmp, tmp, self.rdm1)
                pr = self.rdm1.view(B, d, d).diagonal(dim1=1, dim2=2).real.clone()
                pr = pr / pr.sum(dim=1, keepdim=True)
                cdf = torch.cumsum(pr, dim=1)
                cdf = cdf / cdf[:, -1:]
                idx = torch.searchsorted(cdf, u[:, shot, site:site + 1].contiguous(), right=True).squeeze(1).clamp_(max=d - 1)
                out[:, shot] |= idx << site
                if site != n - 1:
                    sel.zero_()
                    sel[rows, idx] = 1.0                                  # op[0, idx] = 1: slice 0 <- slice idx, the rest <- 0
                    K.apply_local_op(B, None, None, 0, one, er, d, sel.view(-1), d * d, tmp, lay.cap[site + 1])
                    self.scale_buf.copy_(torch.rsqrt(pr[rows, idx]))
                    K.scale(B, None, one, er, d, self.scale_buf, tmp, d * lay.cap[site + 1])
                    K.copy_packed(B, None, one, er, 1, tmp, carry, lay.cap[site + 1])
        return out

    def diagnostics(self, st: BatchedState) -> torch.Tensor:
        """[B, 3]: cost = sum chi^3, max bond (includes d, mps.py:560-565), total bond (mps.py:549-602)."""
        out = torch.zeros((self.B, 3), dtype=torch.float64, device=self.dev)
        self.K.diagnostics(self.B, self.L, self.d, st.bond, out, 3)
        return out

    # ------------------------------------------------------------------------------------------------------------------
    # stochastic process (stochastic_process.py:36-292)
    # ------------------------------------------------------------------------------------------------------------------
    def _build_jump_table(self) -> list[JumpProcess]:
        from .noise import is_pauli
        table: list[JumpProcess] = []
        n = self.L
        for site in range(n):
            for pr in self.noise:
                if len(pr["sites"]) == 1 and pr["sites"][0] == site:
                    m = np.asarray(pr["matrix"], dtype=np.complex128)
                    table.append(JumpProcess("one", (site,), float(pr["strength"]), m.conj().T @ m, (m,)))
            if site < n - 1:
                for pr in self.noise:
                    if len(pr["sites"]) == 2 and pr["sites"][0] == site:
                        if is_pauli(pr):
                            if abs(pr["sites"][1] - site) > 1:
                                ops = tuple(np.asarray(f, dtype=np.complex128) for f in pr["factors"])
                                table.append(JumpProcess("pauli_lr", tuple(pr["sites"]), float(pr["strength"]), None, ops))
                            else:
                                m = np.asarray(pr["matrix"], dtype=np.complex128)
                                table.append(JumpProcess("pauli_adj", tuple(pr["sites"]), float(pr["strength"]), None, (m,)))
                        elif pr["sites"][1] == site + 1:
                            m = np.asarray(pr["matrix"], dtype=np.complex128)
                            table.append(JumpProcess("adjacent", tuple(pr["sites"]), float(pr["strength"]), m.conj().T @ m, (m,)))
                        else:
                            raise NotImplementedError("Non-Pauli long-range two-site jumps are not supported")
        return table

    def _jump_tables(self) -> dict:
        """Device tables of the current noise context's jump table (built once per context): for the weights, L^dagger L stacks
        with site / column indices per kind; for the application, the operator of every process as a 2 x 2 / 4 x 4 row."""
        ctx = self._nctx
        tab = ctx.get("dev")
        if tab is None:
            procs, dev = self.jump_table, self.dev
            eye2, eye4 = np.eye(2, dtype=np.complex128), np.eye(4, dtype=np.complex128)
            one = [(k, pr.sites[0], pr.ldl if pr.kind == "one" else eye2) for k, pr in enumerate(procs) if pr.kind in ("one", "pauli_lr", "pauli_adj")]
            two = [(k, pr.sites[0], pr.ldl) for k, pr in enumerate(procs) if pr.kind == "adjacent"]
            tab = {"need1": sorted({s for _, s, _ in one}), "need2": sorted({s for _, s, _ in two}),
                   "strength": torch.tensor([pr.strength for pr in procs], dtype=torch.float64, device=dev)}
            for name, items in (("one", one), ("two", two)):
                if items:
                    tab[name] = (torch.tensor([k for k, _, _ in items], dtype=torch.int64, device=dev),
                                 torch.tensor([s_ for _, s_, _ in items], dtype=torch.int64, device=dev),
                                 torch.from_numpy(np.stack([m for _, _, m in items]).astype(np.complex128)).to(dev))
            # application tables (host copies drive the per-site launch plan, device copies are gathered by `choice`)
            tab["kind"] = [pr.kind for pr in procs]
            tab["site_a"] = np.array([pr.sites[0] if pr.kind in ("one", "pauli_lr") else -1 for pr in procs], dtype=np.int32)
            tab["site_b"] = np.array([pr.sites[1] if pr.kind == "pauli_lr" else -1 for pr in procs], dtype=np.int32)
            tab["site_c"] = np.array([pr.sites[0] if pr.kind in ("adjacent", "pauli_adj") else -1 for pr in procs], dtype=np.int32)
            tab["ops_a"] = torch.from_numpy(np.stack([pr.ops[0] if pr.kind in ("one", "pauli_lr") else eye2 for pr in procs]).astype(np.complex128)).to(dev)
            tab["ops_b"] = torch.from_numpy(np.stack([pr.ops[1] if pr.kind == "pauli_lr" else eye2 for pr in procs]).astype(np.complex128)).to(dev)
            # a unitary one-site operator (every Pauli) leaves its site a left isometry: no literal merge + split behind it
            def unitary(m):
                return bool(np.array_equal(m.conj().T @ m, np.eye(m.shape[0])))
            tab["unitary_a"] = np.array([unitary(pr.ops[0]) if pr.kind in ("one", "pauli_lr") else True for pr in procs], dtype=bool)
            tab["unitary_b"] = np.array([unitary(pr.ops[1]) if pr.kind == "pauli_lr" else True for pr in procs], dtype=bool)
            tab["ops_c"] = torch.from_numpy(np.stack([pr.ops[0] if pr.kind in ("adjacent", "pauli_adj") else eye4 for pr in procs]).astype(np.complex128)).to(dev)
            # process tables of yb_jump_select: site, one- / two-site weight, L^dagger L (identity for Pauli processes), operator rows
            ldl = np.zeros((len(procs), 16), dtype=np.complex128)
            for k, pr in enumerate(procs):
                if pr.kind == "adjacent":
                    ldl[k] = np.asarray(pr.ldl, dtype=np.complex128).reshape(-1)
                else:
                    ldl[k, :4] = (np.asarray(pr.ldl, dtype=np.complex128) if pr.kind == "one" else eye2).reshape(-1)
            tab["psite"] = torch.tensor([pr.sites[0] for pr in procs], dtype=torch.int32, device=dev)
            tab["ptwo"] = torch.tensor([1 if pr.kind == "adjacent" else 0 for pr in procs], dtype=torch.int32, device=dev)
            tab["pldl"] = torch.from_numpy(ldl).to(dev)
            tab["ops_all"] = torch.cat([tab["ops_a"].reshape(len(procs), 4), tab["ops_b"].reshape(len(procs), 4), tab["ops_c"].reshape(len(procs), 16)],
                                       dim=1).contiguous()
            ctx["dev"] = tab
        return tab

    def stochastic_process(self, st: BatchedState, dt: float, u1, u2: torch.Tensor | None = None, log: list | None = None,
                           tag: str = "") -> torch.Tensor | None:
        """One jump test for every trajectory (stochastic_process.py:190-292).  ``u1``: a ``Uniforms`` source, or (older callers) the
        [B] tensors u1 / u2 themselves (u2 is consumed only by trajectories that jump).  Returns the int32 jump mask (or None when
        there is no noise model).  The test itself (dp = 1 - norm, u1 < dp, stream positions) is yb_jump_test; when somebody jumps,
        one transfer sweep gives the local density matrices and yb_jump_select turns them into weights, pdf, the selected process
        and its operator rows.  Host synchronisations: one read of the jump mask, and -- only when somebody jumps -- one read of
        (weight totals, selected processes), from which the whole launch plan of the jump application is made on the host."""
        K, B, lay, d, n = self.K, self.B, self.lay, self.d, self.L
        K.phase = "jumps"
        if self.noise is None:
            self.normalize_site(st, 0)                      # :226-231 (the uniform is not drawn)
            return None
        src = u1 if isinstance(u1, Uniforms) else Uniforms(table=torch.stack([u1, u2], dim=1).contiguous())
        K.norm2(B, None, st.ext(0), st.ext(1), d, st.site(0), self.norm_buf)
        K.jump_test(B, self.norm_buf, src.table, src.col, src.pos, src.philox, self.jt_dp, self.jt_u1, self.jt_u2, self.jt_mask, self.jt_n)
        jmask = self.jt_mask.clone()                        # (the buffer belongs to the next test)
        jump_h = jmask.cpu().numpy().astype(bool)           # sync 1
        njump = int(jump_h.sum())
        rec = None
        if log is not None:
            rec = dict(tag=tag, dp=self.jt_dp.cpu().numpy().copy(), u1=self.jt_u1.cpu().numpy().copy(), jumped=jump_h.copy(), choice=None, pdf=None)
            log.append(rec)
        if njump < B:
            self.normalize_site(st, 0, None if njump == 0 else (1 - jmask))
        if njump == 0:
            return jmask
        self.stats["jumps"] += njump
        tab = self._jump_tables()
        nproc = len(self.jump_table)
        if self.jt_w is None or self.jt_w.shape[1] != nproc:
            self.jt_w = torch.zeros((B, nproc), dtype=torch.float64, device=self.dev)
        self.rho_sweep(st, tab["need1"], tab["need2"], jmask)          # local density matrices of the jumping trajectories
        K.jump_select(B, jmask, self.jt_u2, self._rdm1_all, self._rdm2_all if tab["need2"] else None, nproc, tab["psite"], tab["ptwo"], tab["pldl"],
                      tab["strength"], dt, tab["ops_all"], self.JT_NOPS, self.jt_w, self.jt_total, self.jt_choice, self.jt_ops, self.status)
        choice = self.jt_choice.to(torch.int64)
        host = torch.stack([self.jt_total, choice.to(torch.float64)]).cpu().numpy()      # sync 2
        total_h, choice_h = host[0], np.where(jump_h, host[1], 0).astype(np.int64)
        bad_h = jump_h & ~(np.isfinite(total_h) & (total_h > 0))
        if bad_h.any():
            raise ValueError("Jump probability weights are zero or non-finite. Reduce process strengths and/or the timestep dt so that "
                             "dt * strength * ||L|psi>||^2 remains representable.")
        if rec is not None:
            rec["choice"] = choice_h.copy()
            rec["pdf"] = self.jt_w.cpu().numpy().copy()
            rec["u2"] = self.jt_u2.cpu().numpy().copy()
        self._apply_jumps(st, jmask, choice, jump_h, choice_h)
        return jmask

    def has_scheduled_jump(self, time: float) -> bool:
        """core/methods/scheduled_jumps.py:29-52."""
        return any(np.isclose(j["time"], time, atol=self.p.dt * 1e-3, rtol=0.0) for j in self.scheduled_jumps)

    def apply_scheduled_jumps(self, st: BatchedState, time: float) -> None:
        """apply_scheduled_jumps (core/methods/scheduled_jumps.py:55-119) for the whole batch: the deterministic jump operators
        of this grid time act on every trajectory (one-site in place, adjacent two-site on the merged tensor followed by the
        run's truncated split), then normalize("B"): a right-to-left sweep of truncation-free shifts and division by the norm,
        which is checked first (a jump that annihilates the state raises, :108-115)."""
        K, B, lay, d, n, p = self.K, self.B, self.lay, self.d, self.L, self.p
        user = self._trunc_args(p.trunc_mode, p.svd_threshold, p.max_bond_dim, 1)
        keep_all = self._trunc_args("discarded_weight", 0.0, None, 1)
        for jump in self.scheduled_jumps:
            if not np.isclose(jump["time"], time, atol=p.dt * 1e-3, rtol=0.0):
                continue
            sites = list(jump["sites"])
            op = self._const(jump["matrix"])
            if len(sites) == 1:
                s_ = sites[0]
                K.apply_local_op(B, None, None, 0, st.ext(s_), st.ext(s_ + 1), d, op, 0, st.site(s_), lay.cap[s_] * lay.cap[s_ + 1])
            else:
                a = min(sites)
                if abs(sites[0] - sites[1]) != 1:
                    raise ValueError(f"Scheduled jump acts on non-adjacent sites {sites}. Only nearest-neighbor jumps are supported.")
                th = self.merge(st, a)
                K.apply_local_op(B, None, None, 0, st.ext(a), st.ext(a + 2), d * d, op, 0, th, lay.cap[a] * lay.cap[a + 2])
                self.split_theta(st, a, "right", user, None, fast=True)
        for i in reversed(range(1, n)):
            self.shift_left(st, i, keep_all, None, fast=True)
        K.norm2(B, None, st.ext(0), st.ext(1), d, st.site(0), self.norm_buf)
        bad = ~(torch.isfinite(self.norm_buf) & (self.norm_buf > 0))
        if K.sync_int(bad.to(torch.int32).sum()) > 0:
            val = float(self.norm_buf[bad][0].item())
            raise ValueError(f"Scheduled jump produced a zero or non-finite squared norm (squared_norm={val}). "
                             "The jump operator annihilates the current state.")
        self.normalize_site(st, 0)

    def _apply_jumps(self, st: BatchedState, jmask: torch.Tensor, choice: torch.Tensor, jump_h: np.ndarray, choice_h: np.ndarray) -> None:
        """Left-canonicalise the jumping trajectories (the reference's PDF sweep leaves the centre at L-1,
        stochastic_process.py:139-141), apply the selected operators (:251-288) and renormalise with the right-to-left
        SVD sweep of normalize("B", "SVD") (:291, mps.py:815-839).  ``jump_h`` / ``choice_h``: host copies of the jump mask and
        the selected process per trajectory -- every launch decision below is made from them, no further host reads."""
        K, B, lay, d, n, p = self.K, self.B, self.lay, self.d, self.L, self.p
        tab = self._jump_tables()
        dev = self.dev
        keep_all = self._trunc_args("discarded_weight", 0.0, None, 1)
        shift = self._trunc_args("discarded_weight", SHIFT_THRESHOLD, None, 1)
        user = self._trunc_args(p.trunc_mode, p.svd_threshold, p.max_bond_dim, 1)
        for i in range(n - 1):
            self.shift_right(st, i, keep_all, jmask, fast=True)
        # per-trajectory sites (host) and operators (device, gathered by the selected process)
        sa_h = np.where(jump_h, tab["site_a"][choice_h], -1).astype(np.int32)
        sb_h = np.where(jump_h, tab["site_b"][choice_h], -1).astype(np.int32)
        sc_h = np.where(jump_h, tab["site_c"][choice_h], -1).astype(np.int32)
        # site whose left-isometry property a jump breaks: the site itself for one-site operators that are not unitary (a Pauli jump
        # leaves A^dagger A = 1 untouched, so the reduced shift behind it is still the reference's merged SVD), the right site of a pair
        broken = np.zeros((n, B), dtype=bool)
        safe_choice = np.where(jump_h, choice_h, 0)
        for arr, off, keeps in ((sa_h, 0, tab["unitary_a"][safe_choice]), (sb_h, 0, tab["unitary_b"][safe_choice]), (sc_h, 1, None)):
            rows = np.nonzero((arr >= 0) if keeps is None else ((arr >= 0) & ~keeps))[0]
            broken[arr[rows] + off, rows] = True
        lit_h = np.zeros((n, B), dtype=np.int32)           # row i: trajectories that need the literal merge + split at bond i
        lit_h[1:] = (broken[:-1] & jump_h[None, :])
        plan = torch.from_numpy(np.concatenate([sa_h[None], sb_h[None], sc_h[None], lit_h])).to(dev)      # one upload
        site_a, site_b, site_c = plan[0], plan[1], plan[2]
        lit_d = plan[3: 3 + n]
        # the operator rows of the selected processes were gathered by yb_jump_select: [B][4 | 4 | 16] = first-site op, second-site op, pair op
        nops = self.JT_NOPS
        ops_a, ops_b, ops_c = self.jt_ops, self.jt_ops[4:], self.jt_ops[8:]
        if (sa_h >= 0).any():
            mask_a = (site_a >= 0).to(torch.int32)
            for s_ in np.unique(sa_h[sa_h >= 0]):
                s_ = int(s_)
                K.apply_local_op(B, mask_a, site_a, s_, st.ext(s_), st.ext(s_ + 1), d, ops_a, nops, st.site(s_), lay.cap[s_] * lay.cap[s_ + 1])
        if (sb_h >= 0).any():
            mask_b = (site_b >= 0).to(torch.int32)
            for s_ in np.unique(sb_h[sb_h >= 0]):
                s_ = int(s_)
                K.apply_local_op(B, mask_b, site_b, s_, st.ext(s_), st.ext(s_ + 1), d, ops_b, nops, st.site(s_), lay.cap[s_] * lay.cap[s_ + 1])
        if (sc_h >= 0).any():
            for s_ in np.unique(sc_h[sc_h >= 0]):
                s_ = int(s_)
                m_s = (site_c == s_).to(torch.int32)
                th = self.merge(st, s_, m_s)
                K.apply_local_op(B, m_s, None, 0, st.ext(s_), st.ext(s_ + 2), d * d, ops_c, nops, th, lay.cap[s_] * lay.cap[s_ + 2])
                self.split_theta(st, s_, "right", user, m_s, fast=True)
        # normalize("B", "SVD"): right-to-left merged SVD shifts (mps.py:815-839).  Where the left neighbour is still a left isometry the
        # reduced shift has the merged pair's singular values; where it carries a non-unitary jump operator it is made an isometry again
        # first -- A'_{i-1} = Q R without truncation (a keep-all shift to the right: Cholesky-QR, or the eigen-solver route when the
        # operator left it rank-deficient), R into the centre -- so that theta = Q (R C_i) has the singular values of R C_i and the
        # reduced shift decides exactly what the reference's merged SVD decides.  (Round 1 took the literal merge + split: a 256 x 256
        # eigen-solve for the one or two trajectories concerned, 3.9 ms of latency each.)
        nlit = lit_h.sum(axis=1)
        for i in reversed(range(1, n)):
            if nlit[i] > 0:
                self.shift_right(st, i - 1, keep_all, lit_d[i], fast=True)
            self.shift_left(st, i, shift, jmask, fast=True)
        self.normalize_site(st, 0, jmask)

    # ------------------------------------------------------------------------------------------------------------------
    # trajectory drivers (analog/analog_tjm.py:86-462)
    # ------------------------------------------------------------------------------------------------------------------
    def run(self, traj_uniforms: torch.Tensor | None, sample_uniforms: torch.Tensor | None, log: list | None = None, **kwargs):
        """All time steps in one call: exhausts ``run_iter`` (same arguments) and returns (results [B, n_obs, T | 1],
        diagnostics [B, 3, T | 1]) as device tensors."""
        for _ in self.run_iter(traj_uniforms, sample_uniforms, log, **kwargs):
            pass
        return self.results, self.diag

    def run_iter(self, traj_uniforms: torch.Tensor | None, sample_uniforms: torch.Tensor | None, log: list | None = None, *,
                 start_pos: torch.Tensor | None = None, sample_at: Sequence[int] | None = None, continue_trajectory: bool = False,
                 interval_mpos: Sequence[Sequence[np.ndarray]] | None = None, final_sample_from_trajectory_stream: bool = False):
        """Generator over the time grid: yields the time index j after ALL work belonging to it has been enqueued (for order 2: the
        trajectory step and its measurement copy), so a caller can stream that step's inputs in before resuming and its result
        column (``self.results[:, :, col]``, ``self.diag``) out afterwards -- ``Simulator.open`` does exactly that.  Uniforms are
        read when a step runs: columns 2j, 2j+1 of the trajectory table and row j of the sample table must be on the device before
        the generator is resumed for time index j (a jump test at step j reads a position <= 2j - 1).
        Run all time steps.  traj_uniforms: [B, >= 2*n_times] trajectory stream (consumed 1 or 2 per jump test);
        sample_uniforms: [B, n_times, 2] per-timestep measurement streams (order 2).  Returns (results [B, n_obs, T],
        diagnostics [B, 3, T]) as device tensors.

        Keyword arguments of SimulationProgram execution (analog_tjm.py:206-366,369-462; simulator.py:339-552):
        ``start_pos`` [B] int64 -- where each trajectory's stream continues (one Generator is shared by all segments of a program;
        ``self._pos`` holds the position after the run); ``sample_at`` -- time-grid indices to record instead of the
        ``sample_timesteps`` rule (:50-62); ``continue_trajectory`` -- order 2 only: the loaded state is a mid-Trotter phi, the
        junction is re-measured and ``initialize`` is skipped (:312-324); ``interval_mpos`` -- one MPO per interval of the grid
        (merged segments with different Hamiltonians): step j evolves with interval j-1 (order 1), with j-2 for step_through and
        j-1 for its measurement copy (order 2, :343-360); ``final_sample_from_trajectory_stream`` -- order 2: the measurement copy
        of the last time draws from the trajectory stream instead of its own sample stream (:271-279)."""
        p = self.p
        nt = p.n_times
        ncol = nt if p.sample_timesteps else 1
        nobs = len(self.observables)
        results = torch.zeros((self.B, nobs, ncol), dtype=torch.float64, device=self.dev)
        diag = torch.zeros((self.B, 3, ncol), dtype=torch.float64, device=self.dev)
        self.results, self.diag = results, diag
        self.spectra: dict = {}                             # (observable row, column) -> [B, 500] Schmidt spectra
        self._pos = torch.zeros(self.B, dtype=torch.int64, device=self.dev) if start_pos is None else start_pos.to(self.dev, torch.int64).clone()
        self._traj_u = traj_uniforms
        measure_at = None
        if sample_at is not None:                           # _validate_sample_at (:65-83)
            for index in sample_at:
                if isinstance(index, bool) or not isinstance(index, (int, np.integer)) or index < 0 or index >= nt:
                    raise ValueError(f"sample_at index {index!r} is outside the time grid [0, {nt}).")
            measure_at = frozenset(int(i) for i in sample_at)
            if not p.sample_timesteps and len(measure_at) > 1:
                raise ValueError("Selecting multiple sample_at indices requires sample_timesteps=True.")
        if interval_mpos is not None and len(interval_mpos) != nt - 1:
            raise ValueError(f"interval_mpos needs one MPO per interval ({nt - 1}), got {len(interval_mpos)}")

        def wanted(j: int) -> bool:                         # _measure_at (:50-62)
            if measure_at is not None:
                return j in measure_at
            return p.sample_timesteps or j == nt - 1

        def use_interval(k: int) -> None:
            if interval_mpos is not None:
                key = id(interval_mpos[k])
                if key not in self._mpo_cache:
                    self._mpo_cache[key] = self._upload_mpo(interval_mpos[k])
                self._set_mpo(self._mpo_cache[key])

        def record(state, col):
            diag[:, :, col] = self.diagnostics(state)
            if nobs:
                results[:, :, col] = self.evaluate_observables(state)
                for k, spec in self._spectra_now.items():
                    self.spectra[(k, col)] = spec

        def traj_jump_test(state, tag):
            if self.noise is None:
                self.stochastic_process(state, p.dt, None, None, log, tag)
                return
            # the trajectory stream: position _pos[b], advanced inside yb_jump_test by the doubles the reference's Generator hands out
            if self.philox is not None:
                src = Uniforms(philox=(self.philox[0], self.philox[1], STREAM_TRAJECTORY, 0), pos=self._pos)
            else:
                src = Uniforms(table=self._traj_u, pos=self._pos)
            self.stochastic_process(state, p.dt, src, None, log, tag)

        phi = self.phi
        if p.order == 2:
            psi = self.psi

            def sample(j, interval):
                if not wanted(j):                           # the last column is always wanted on the paths the callers use
                    return
                use_interval(interval)
                psi.copy_from(phi)
                self.tdvp(psi)
                self.apply_dissipation(psi, p.dt / 2)
                if self.noise is None:
                    self.stochastic_process(psi, p.dt, None, None, log, f"sample{j}")
                elif final_sample_from_trajectory_stream and j == nt - 1:
                    traj_jump_test(psi, f"sample{j}")
                elif self.philox is not None:               # the measurement copy's own stream: block j + offset of the "SAMP" stream
                    self.stochastic_process(psi, p.dt, Uniforms(philox=(self.philox[0], self.philox[1], STREAM_SAMPLE, j + self.philox[2])), None,
                                            log, f"sample{j}")
                else:
                    self.stochastic_process(psi, p.dt, Uniforms(table=sample_uniforms.view(self.B, -1), col=2 * j), None, log, f"sample{j}")
                record(psi, j if p.sample_timesteps else 0)

            if nt == 1:
                if wanted(0):
                    record(phi, 0)
                yield 0
                return
            if continue_trajectory:                         # :312-324 (static operator: interval 0 everywhere)
                sample(0, 0)
                yield 0
                for j in range(1, nt):
                    use_interval(0)
                    self.tdvp(phi)
                    self.apply_dissipation(phi, p.dt)
                    traj_jump_test(phi, f"step{j}")
                    sample(j, 0)
                    yield j
                return
            if measure_at is not None and 0 in measure_at or measure_at is None and p.sample_timesteps:
                record(phi, 0)
            yield 0
            self.apply_dissipation(phi, p.dt / 2)           # initialize (:86-108)
            traj_jump_test(phi, "init")
            sample(1, 0)
            yield 1
            for j in range(2, nt):
                use_interval(j - 2)
                self.tdvp(phi)                              # step_through (:111-140)
                self.apply_dissipation(phi, p.dt)
                traj_jump_test(phi, f"step{j}")
                sample(j, j - 1)
                yield j
            return
        # first-order scheme (:369-462)
        if continue_trajectory:
            raise ValueError("continue_trajectory is an order-2 hand-off")
        if self.has_scheduled_jump(0.0):                    # :428-431: before the initial sample
            self.apply_scheduled_jumps(phi, 0.0)
        measure_initial = (0 in measure_at) if measure_at is not None else p.sample_timesteps
        if measure_initial:
            record(phi, 0)
        if nt > 1:
            yield 0
        for j in range(1, nt):
            use_interval(j - 1)
            self.tdvp(phi)
            if self.noise is not None or self.scheduled_jumps:
                if self.noise is not None:
                    self.apply_dissipation(phi, p.dt)
                if self.has_scheduled_jump(j * p.dt):       # replaces the jump test of this step (no uniform is drawn)
                    self.apply_scheduled_jumps(phi, j * p.dt)
                else:
                    traj_jump_test(phi, f"step{j}")
            if wanted(j):
                record(phi, j if p.sample_timesteps else 0)
            yield j
        if not measure_initial and not p.sample_timesteps and nt <= 1:
            record(phi, 0)
        if nt <= 1:
            yield 0

    def steady_state_step(self, u_traj: torch.Tensor, u_sample: torch.Tensor):
        """One steady-state sampled time step of the second-order scheme for the whole batch: step_through on the
        trajectory state (TDVP sweep, dissipation dt, jump test) followed by sample() on a copy (TDVP sweep, dissipation
        dt/2, jump test, diagnostics + observables) -- analog_tjm.py:111-199,352-360.  u_traj / u_sample: [B, 2] uniforms.
        Returns (observables [B, n_obs], diagnostics [B, 3]).  Used by bench.py to time exactly one unit of work."""
        p, phi, psi = self.p, self.phi, self.psi
        self.tdvp(phi)
        self.apply_dissipation(phi, p.dt)
        self.stochastic_process(phi, p.dt, u_traj[:, 0].contiguous(), u_traj[:, 1].contiguous())
        psi.copy_from(phi)
        self.tdvp(psi)
        self.apply_dissipation(psi, p.dt / 2)
        self.stochastic_process(psi, p.dt, u_sample[:, 0].contiguous(), u_sample[:, 1].contiguous())
        return self.evaluate_observables(psi), self.diagnostics(psi)

    # ------------------------------------------------------------------------------------------------------------------
    # digital TJM (digital/digital_tjm.py:187-749): gates, per-gate local noise, TDVP gate windows
    # ------------------------------------------------------------------------------------------------------------------
    def apply_single_qubit_gate(self, st: BatchedState, gate) -> None:
        """_apply_single_qubit_gate (digital_tjm.py:304-309).  A unitary on the physical index keeps every isometry
        condition, so the batch stays in B form (the reference merely loses track of the centre)."""
        s = gate.sites[0]
        op = self._const(gate.matrix)
        self.K.apply_local_op(self.B, None, None, 0, st.ext(s), st.ext(s + 1), self.d, op, 0, st.site(s), self.lay.cap[s] * self.lay.cap[s + 1])

    @staticmethod
    def _resolve_lr_matrix(gate) -> np.ndarray:
        """resolve_lr_tensor / convert_nn_matrix (mpo_utils.py:104-160) as a 4x4 on the merged index s_left*2 + s_right."""
        m = np.asarray(gate.matrix, dtype=np.complex128)
        if gate.sites[0] < gate.sites[1]:
            return m
        t = np.zeros((2, 2, 2, 2), dtype=np.complex128)
        for col in range(4):
            il, ir = col % 2, col // 2
            for row in range(4):
                ol, orr = row % 2, row // 2
                t[ol, orr, il, ir] = m[row, col]
        return t.reshape(4, 4)

    def _move_centre(self, st: BatchedState, cur: int, target: int) -> None:
        """shift_center_to (mps.py:374-393) with truncation-free shifts (QR-equivalent)."""
        keep_all = self._trunc_args("discarded_weight", 0.0, None, 1)
        for i in range(cur, target):
            self.shift_right(st, i, keep_all, fast=True)
        for i in range(cur, target, -1):
            self.shift_left(st, i, keep_all, fast=True)

    def apply_two_qubit_gate_tebd(self, st: BatchedState, gate, centre: int = 0) -> int:
        """apply_two_qubit_gate_tebd (digital_tjm.py:455-533), entered with the orthogonality centre at ``centre``.  Nearest
        neighbours: truncation-free shifts bring the centre onto the pair, merge, gate, truncated split with S to the right.
        Other pairs go through the SWAP ladder of :463-481 (swaps from the right end down to left+1, the gate on
        (left, left+1), swaps back).  Returns the new centre (the right site of the last pair touched)."""
        lay, K, B, d, p = self.lay, self.K, self.B, self.d, self.p
        s0, s1 = gate.sites
        if abs(s0 - s1) != 1:
            left, right = min(s0, s1), max(s0, s1)

            def swap_at(i: int, c: int) -> int:
                g = GateLibrary.swap()
                g.set_sites(i, i + 1)
                return self.apply_two_qubit_gate_tebd(st, g, c)

            for i in range(right - 1, left, -1):
                centre = swap_at(i, centre)
            adj = copy.deepcopy(gate)
            adj.set_sites(*((left, left + 1) if s0 == left else (left + 1, left)))
            centre = self.apply_two_qubit_gate_tebd(st, adj, centre)
            for i in range(left + 1, right):
                centre = swap_at(i, centre)
            return centre
        left = min(s0, s1)
        if centre not in (left, left + 1):
            self._move_centre(st, centre, left)
        th = self.merge(st, left)
        op = self._const(self._resolve_lr_matrix(gate))
        K.apply_local_op(B, None, None, 0, st.ext(left), st.ext(left + 2), d * d, op, 0, th, lay.cap[left] * lay.cap[left + 2])
        trunc = self._trunc_args(p.trunc_mode, p.svd_threshold, p.max_bond_dim, p.tdvp_min_keep)
        self.split_theta(st, left, "right", trunc, fast=True)
        return left + 1

    @staticmethod
    def gate_support_mpo(gate) -> list[np.ndarray]:
        """BaseGate.set_sites -> extend_gate -> split_tensor (gate_library.py:25-126,196-198; what get_support_mpo returns from the
        gate's cache, mpo_utils.py:198-204): the gate tensor with its axes permuted to ascending site order, legs grouped per site,
        split site by site with SVDs that carry the singular values to the right under the reference's hard cut-off 1e-6
        (min_keep 1); identity tensors of the carried bond on spectator sites.  Returns the (2, 2, Dl, Dr) tensors for the sites
        min(sites) .. max(sites).  Two- and three-qubit gates."""
        n = int(gate.interaction)
        t = np.asarray(gate.matrix, dtype=np.complex128).reshape((2,) * (2 * n))
        order = sorted(range(n), key=lambda k: gate.sites[k])
        t = np.transpose(t, [*order, *[n + k for k in order]])
        srt = sorted(gate.sites)
        rem = np.transpose(t, [ax for k in range(n) for ax in (k, n + k)]).reshape(4, 4 ** (n - 1))
        parts, left = [], 1
        for _ in range(n - 1):
            u, sv, vh = np.linalg.svd(rem, full_matrices=False)
            keep = max(int(np.sum(sv > 1e-6)), 1)
            parts.append(np.transpose(u[:, :keep].reshape(left, 2, 2, keep), (1, 2, 0, 3)))
            left = keep
            rem = (np.diag(sv[:keep]) @ vh[:keep]).reshape(left * 4, rem.shape[1] // 4)
        parts.append(np.transpose(rem.reshape(left, 2, 2), (1, 2, 0))[:, :, :, None])
        tensors = [parts[0]]
        for k in range(1, n):
            bond = tensors[-1].shape[3]
            for _ in range(srt[k - 1] + 1, srt[k]):
                ident = np.zeros((2, 2, bond, bond), dtype=np.complex128)
                for q in range(bond):
                    ident[:, :, q, q] = np.eye(2)
                tensors.append(ident)
            tensors.append(parts[k])
        return [np.ascontiguousarray(w, dtype=np.complex128) for w in tensors]

    @staticmethod
    def mpo_route_bond_factor(layers, gate_mode: str, length: int) -> list[int] | None:
        """Per-bond capacity multiplier a circuit needs on the gate-MPO route (None when no gate takes it): the operator-Schmidt
        rank of every non-adjacent two-qubit gate that ``run_digital`` will send through ``apply_long_range_gate_mpo``."""
        factor = [1] * (length + 1)
        for layer in layers:
            for gate in (*layer.even_two_qubit_gates, *layer.odd_two_qubit_gates):
                if gate.interaction == 2 and (abs(gate.sites[0] - gate.sites[1]) == 1 or gate_mode == "swaps"):
                    continue
                if gate_mode in ("tdvp", "full-tdvp") and getattr(gate, "generator", None) is not None:
                    continue
                for k, w in enumerate(BatchedTJM.gate_support_mpo(gate)[:-1]):
                    j = min(gate.sites) + 1 + k
                    factor[j] = max(factor[j], w.shape[3])
        return factor if max(factor) > 1 else None

    def apply_long_range_gate_mpo(self, st: BatchedState, gate) -> int:
        """apply_long_range_gate_mpo (digital_tjm.py:536-556): ``MPO.from_gate(gate).multiply(state, compress=True)`` for the whole
        batch.  (1) every support site A <- W.A with fused bonds (yb_mpo_site_apply; sites outside the support see the identity
        and are left alone); (2) ``MPS.compress`` (mps.py:841-899): right-canonical gauge without truncation, one left-to-right
        sweep of merged two-site SVDs under the user's truncation policy (min_keep 1), centre moved back without truncation.
        The reference restores the centre ``check_canonical_form`` finds on the un-gauged product, else L // 2; a gate's
        operator-Schmidt factors are not isometries on both of its sites, so the fallback is what it takes -- used here for
        every trajectory (the choice is a gauge, it does not enter any observable).  Returns the new centre."""
        lay, K, B, d, n, p = self.lay, self.K, self.B, self.d, self.L, self.p
        first, last = min(gate.sites), max(gate.sites)
        key = ("gate-mpo", gate.name, tuple(gate.sites), np.asarray(gate.matrix, dtype=np.complex128).tobytes())
        if key not in self._mpo_cache:
            self._mpo_cache[key] = [(torch.from_numpy(w.reshape(-1)).to(self.dev), w.shape[2], w.shape[3]) for w in self.gate_support_mpo(gate)]
        ws = self._mpo_cache[key]
        ranks = torch.tensor([dr for _, _, dr in ws[:-1]], dtype=st.bond.dtype, device=self.dev)      # MPO bond on bonds first+1 .. last
        old = st.bond.clone()
        grown = old[:, first + 1: last + 1] * ranks[None, :]
        if bool((grown.max(dim=0).values.cpu() > torch.tensor(lay.cap[first + 1: last + 1])).any()):
            raise ValueError("gate-MPO route: chi * D_gate exceeds the slot capacity; build the engine with bond_factor="
                             "BatchedTJM.mpo_route_bond_factor(layers, gate_mode, length)")
        st.bond[:, first + 1: last + 1] = grown.to(st.bond.dtype)
        tmp = self._buf(self.tmp, lay.nloc_max)
        for s in range(first, last + 1):
            w, dl, dr = ws[s - first]
            K.mpo_site_apply(B, None, Ext(old, s, n + 1), Ext(old, s + 1, n + 1), d, dl, dr, w, st.site(s), tmp, lay.cap[s] * lay.cap[s + 1])
            K.copy_packed(B, None, st.ext(s), st.ext(s + 1), d, tmp, st.site(s), lay.slot[s])
        keep_all = self._trunc_args("discarded_weight", 0.0, None, 1)
        for i in reversed(range(1, n)):                     # set_canonical_form(0, "QR")
            self.shift_left(st, i, keep_all, fast=True)
        user = self._trunc_args(p.trunc_mode, p.svd_threshold, p.max_bond_dim, 1)
        for i in range(n - 1):
            self.merge(st, i)
            self.split_theta(st, i, "right", user, fast=True)
        target = n // 2
        for i in range(n - 1, target, -1):
            self.shift_left(st, i, keep_all, fast=True)
        return target

    def apply_two_qubit_gate_tdvp(self, st: BatchedState, gate) -> int:
        """apply_two_qubit_gate_tdvp (digital_tjm.py:324-452), entered with the centre at 0: generator MPO of bond
        dimension 1, window = support +- 1 site, one symmetric 2TDVP sweep over the window with evolution time 1.
        Returns the new centre (the first window site)."""
        n, lay, K, B, d, p = self.L, self.lay, self.K, self.B, self.d, self.p
        if p.tdvp_mode != "2site":                   # raised where the reference raises it: only a gate that takes the window route
            raise ValueError(f'apply_two_qubit_gate_tdvp only supports tdvp_mode="2site"; got {p.tdvp_mode!r}.')
        if gate.generator is None:
            raise NotImplementedError("gates without a product-form generator need the gate-MPO route (SURVEY.md 8f)")
        key = (gate.name, tuple(gate.sites), tuple(np.asarray(f).tobytes() for f in gate.generator))
        if key not in self._mpo_cache:
            factors = dict(zip(gate.sites, gate.generator))
            mpo = []
            for s in range(n):
                w = np.zeros((2, 2, 1, 1), dtype=np.complex128)
                w[:, :, 0, 0] = factors.get(s, np.eye(2))
                mpo.append(w)
            self._mpo_cache[key] = self._upload_mpo(mpo)
        first, last = min(gate.sites), max(gate.sites)
        w0, w1 = max(first - 1, 0), min(last + 1, n - 1)
        keep_all = self._trunc_args("discarded_weight", 0.0, None, 1)
        for i in range(w0):                                             # apply_window: centre -> window start (QR)
            self.shift_right(st, i, keep_all, fast=True)
        saved = dict(W1=self.W1, W1T=self.W1T, W2=self.W2)
        self._set_mpo(self._mpo_cache[key])
        try:
            self.sweep_2site(st, [1.0 / p.tdvp_sweeps] * p.tdvp_sweeps, w0, w1)
        finally:
            self._set_mpo(saved)
        if p.max_bond_dim is not None:
            # renorm_drift (sweep_utils.py:231-243): renormalise trajectories whose norm drifted beyond the tolerance.
            # The centre is at w0, so the norm is that of the centre tensor and dividing it out IS the normalisation.
            tol = max(1e-10, float(np.sqrt(p.svd_threshold)))
            K.norm2(B, None, st.ext(w0), st.ext(w0 + 1), d, st.site(w0), self.norm_buf)
            nrm = torch.sqrt(torch.clamp(self.norm_buf, min=0.0))
            self.scale_buf.copy_(torch.where((nrm - 1.0).abs() > tol, 1.0 / nrm, torch.ones_like(nrm)))
            K.scale(B, None, st.ext(w0), st.ext(w0 + 1), d, self.scale_buf, st.site(w0), lay.slot[w0])
        return w0

    def _local_noise_ctx(self, sites) -> dict:
        """create_local_noise_model (digital_tjm.py:187-204): processes whose support lies inside the gate's qubits."""
        key = tuple(sorted(sites))
        if key not in self._noise_cache:
            gs = set(sites)
            self._noise_cache[key] = self._make_noise_ctx([pr for pr in self._global_noise if set(pr["sites"]).issubset(gs)])
        return self._noise_cache[key]

    def run_digital(self, layers, traj_uniforms: torch.Tensor | None, sample_layers: bool, gate_mode: str = "mpo", log: list | None = None,
                    start_pos: torch.Tensor | None = None):
        """digital_tjm (digital_tjm.py:636-749), observable path, for the whole batch.  ``layers``: compiled circuit layers
        (single-qubit gates, even / odd two-qubit groups, sample points); traj_uniforms: [B, 2 * n_two_qubit_gates].
        Returns (results [B, n_obs, cols], diagnostics [B, 3, cols])."""
        p = self.p
        if not p.digital:
            raise ValueError("run_digital needs EngineParams(digital=True)")
        st = self.phi
        glob = self._nctx
        self._global_noise = glob["noise"] or []
        noisy = glob["active"]
        mid = sum(l.sample_points for l in layers) if sample_layers else 0
        ncol = mid + 2 if sample_layers else 1
        nobs = len(self.observables)
        results = torch.zeros((self.B, nobs, ncol), dtype=torch.float64, device=self.dev)
        diag = torch.zeros((self.B, 3, ncol), dtype=torch.float64, device=self.dev)
        pos = torch.zeros(self.B, dtype=torch.int64, device=self.dev) if start_pos is None else start_pos.to(self.dev, torch.int64).clone()
        self._pos = pos                                     # read back by SimulationProgram execution (one stream per trajectory)

        def record(col: int) -> None:
            diag[:, :, col] = self.diagnostics(st)
            if nobs:
                results[:, :, col] = self.evaluate_observables(st)
            if self.on_sample is not None:                  # step-wise consumers (bench.py times Trotter steps between sample points)
                self.on_sample(self, col, results, diag)

        if sample_layers:
            record(0)
        col = 0
        try:
            for layer in layers:
                for gate in layer.single_qubit_gates:
                    self.apply_single_qubit_gate(st, gate)
                for group in (layer.even_two_qubit_gates, layer.odd_two_qubit_gates):
                    for gate in group:
                        has_gen = getattr(gate, "generator", None) is not None
                        nn = gate.interaction == 2 and abs(gate.sites[0] - gate.sites[1]) == 1
                        if gate.interaction > 2:            # no TEBD path: generator window in the TDVP modes, else the gate MPO (:583-590)
                            if gate_mode in ("tdvp", "full-tdvp") and has_gen:
                                centre = self.apply_two_qubit_gate_tdvp(st, gate)
                            else:
                                centre = self.apply_long_range_gate_mpo(st, gate)
                        elif gate_mode == "full-tdvp" and has_gen:
                            centre = self.apply_two_qubit_gate_tdvp(st, gate)
                        elif nn or gate_mode == "swaps":
                            centre = self.apply_two_qubit_gate_tebd(st, gate)
                        elif gate_mode == "tdvp" and has_gen:
                            centre = self.apply_two_qubit_gate_tdvp(st, gate)
                        else:                                   # gate_mode "mpo", or a generator-free gate in a TDVP mode
                            centre = self.apply_long_range_gate_mpo(st, gate)
                        if not noisy:
                            keep_all = self._trunc_args("discarded_weight", 0.0, None, 1)     # normalize("B", "QR")
                            for i in reversed(range(1, centre + 1)):
                                self.shift_left(st, i, keep_all, fast=True)
                            self.normalize_site(st, 0)
                        else:
                            self._set_noise(self._local_noise_ctx(gate.sites))
                            self.apply_dissipation(st, 1.0, start=centre)
                            if self.noise is None:
                                self.stochastic_process(st, 1.0, None, None, log, f"gate{tuple(gate.sites)}")
                            else:
                                if self.philox is not None:
                                    src = Uniforms(philox=(self.philox[0], self.philox[1], STREAM_TRAJECTORY, 0), pos=pos)
                                else:
                                    src = Uniforms(table=traj_uniforms, pos=pos)
                                self.stochastic_process(st, 1.0, src, None, log, f"gate{tuple(gate.sites)}")      # (advances pos)
                if sample_layers:
                    for _ in range(layer.sample_points):
                        col += 1
                        record(col)
        finally:
            self._set_noise(glob)
        record(ncol - 1)
        return results, diag
