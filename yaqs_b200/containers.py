"""Host-side containers that materialise the inputs of a run: MPS, MPO, State, Hamiltonian, Result.

These are the drop-in counterparts of the reference's data structures for ``representation="mps"``:
  MPS          core/data_structures/mps.py:54-301,409-452,815-839   (tensors (s, l, r), presets, pad, normalize)
  MPO          core/data_structures/mpo.py:41-50,248-404,1146-1170  (tensors (s_out, s_in, l, r); ising / heisenberg / pauli / custom)
  State        core/data_structures/state.py:35-60
  Hamiltonian  core/data_structures/hamiltonian.py:36-441
  Result       core/data_structures/result.py:155-189
They run once per ``Simulator.run`` on the host (numpy); the device only consumes their tensors.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

C = np.complex128
_PAULI = {
    "I": np.eye(2, dtype=C), "X": np.array([[0, 1], [1, 0]], dtype=C),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=C), "Z": np.array([[1, 0], [0, -1]], dtype=C),
}
_PRESETS = {
    "zeros": [1, 0], "ones": [0, 1],
    "x+": [1 / np.sqrt(2), 1 / np.sqrt(2)], "x-": [1 / np.sqrt(2), -1 / np.sqrt(2)],
    "y+": [1 / np.sqrt(2), 1j / np.sqrt(2)], "y-": [1 / np.sqrt(2), -1j / np.sqrt(2)],
}


class MPS:
    """Matrix product state on the host: list of (s, l, r) complex128 tensors + tracked orthogonality centre."""

    def __init__(self, length: int, tensors=None, physical_dimensions=None, state: str = "zeros", pad: int | None = None,
                 basis_string: str | None = None) -> None:
        if length <= 0:
            raise ValueError("length must be a positive integer.")
        self.length = length
        dims = physical_dimensions if physical_dimensions is not None else [2] * length
        if isinstance(dims, int):
            dims = [dims] * length
        if any(d != 2 for d in dims):
            raise NotImplementedError("the B200 TJM path supports qubits (d = 2) only")
        self.physical_dimensions = list(dims)
        self.flipped = False
        self.orthogonality_center: int | None
        if tensors is not None:
            if len(tensors) != length:
                raise ValueError("Number of tensors must match the chain length.")
            self.tensors = [np.asarray(t, dtype=C) for t in tensors]
            self.orthogonality_center = None
            # the reference keeps user tensors as given (mps.py:112-114); the batched engine needs them right-canonical with the centre
            # at site 0, so they are re-gauged WITHOUT rescaling: an unnormalised input stays unnormalised, and an evolved state handed
            # back as output_state keeps the norm it has (truncation loss stays visible)
            self._canonicalize(rescale=False)
            return
        self.tensors = []
        if state == "haar-random":
            # mps.py:128-221: a chain of Haar-random isometries, bonds min(pad, 2^k) from either end.  Site by site, left to right, from
            # ONE generator: a complex Ginibre block ((d chi_l) x chi_r: real part drawn first, then the imaginary part), its reduced QR,
            # the column phases of Q fixed so that R has a positive diagonal.  The chain is left-canonical, hence normalised; the engine
            # wants the centre at site 0, so it is re-gauged without rescaling -- the represented vector is the reference's, entry by entry.
            chi = 1 if pad is None else pad
            rng = np.random.default_rng()
            bonds = [1] + [min(chi, 2 ** min(i + 1, length - 1 - i)) for i in range(length - 1)] + [1]
            for i in range(length):
                rows, cols = 2 * bonds[i], bonds[i + 1]
                ginibre = rng.standard_normal((rows, cols)) + 1j * rng.standard_normal((rows, cols))
                q, r = np.linalg.qr(ginibre, mode="reduced")
                diag = np.diag(r)
                phase = np.where(np.abs(diag) > 0, diag / np.where(np.abs(diag) > 0, np.abs(diag), 1.0), 1.0)
                self.tensors.append((q / phase[None, :]).reshape(2, bonds[i], cols).astype(C))
            self.orthogonality_center = None
            self._canonicalize(rescale=False)
            return
        if state == "basis":
            if basis_string is None or len(basis_string) != length:
                raise ValueError("basis_string must be provided with one character per site for 'basis' initialization.")
        for i in range(length):
            if state in _PRESETS:
                v = _PRESETS[state]
            elif state == "Neel":
                v = [1, 0] if i % 2 else [0, 1]
            elif state == "wall":
                v = [1, 0] if i < length // 2 else [0, 1]
            elif state == "basis":
                v = [1, 0] if basis_string[i] == "0" else [0, 1]
            elif state == "random":
                r = np.random.default_rng().random()
                v = [r, 1 - r]
            else:
                raise ValueError("Invalid state string")
            self.tensors.append(np.asarray(v, dtype=C).reshape(2, 1, 1))
        self.orthogonality_center = 0
        if state == "random":
            self.normalize()
        if pad is not None:
            self.pad_bond_dimension(pad)

    # mps.py:815-839 with QR: right-canonical form, centre 0, unit norm
    def normalize(self) -> None:
        """Right-canonical form with the centre at site 0 and unit norm (mps.py:815-839, form "B")."""
        self._canonicalize(rescale=True)

    def _canonicalize(self, rescale: bool) -> None:
        t = self.tensors
        for i in reversed(range(1, self.length)):
            d, l, r = t[i].shape
            q, rm = np.linalg.qr(np.ascontiguousarray(t[i].transpose(0, 2, 1)).reshape(d * r, l))
            t[i] = q.reshape(d, r, q.shape[1]).transpose(0, 2, 1)
            t[i - 1] = np.einsum("slj,kj->slk", t[i - 1], rm)
        nrm = np.linalg.norm(t[0])
        if rescale and nrm > 0:
            t[0] = t[0] / nrm
        self.orthogonality_center = 0

    def pad_bond_dimension(self, target_dim: int) -> None:
        """mps.py:409-452."""
        n = self.length
        for i, t in enumerate(self.tensors):
            d, l, r = t.shape
            lt = 1 if i == 0 else min(target_dim, 2 ** min(i, n - i))
            rt = 1 if i == n - 1 else min(target_dim, 2 ** min(i + 1, n - 1 - i))
            if l > lt or r > rt:
                raise ValueError("Target bond dim must be at least current bond dim.")
            new = np.zeros((d, lt, rt), dtype=C)
            new[:, :l, :r] = t
            self.tensors[i] = new
        self.normalize()

    def to_vector(self) -> np.ndarray:
        acc = self.tensors[0][:, 0, :]
        for t in self.tensors[1:]:
            acc = np.einsum("sr,prq->spq", acc, t).reshape(-1, t.shape[2])
        return acc[:, 0]

    # ------------------------------------------------------------------------------------------------------------------------------
    # Read-only inspection of a host MPS: what a user does with ``Result.output_state.mps`` (mps.py:514-678, 901-1047, 1236-1280,
    # 1495-1660).  Host numpy, transfer-matrix contractions; nothing here is on the device path.
    # ------------------------------------------------------------------------------------------------------------------------------
    def bond_dimensions(self) -> list[int]:
        """Outgoing bond of every site but the last (mps.py:514-520)."""
        return [int(t.shape[2]) for t in self.tensors[:-1]]

    def get_max_bond(self) -> int:
        """mps.py:549-565 (the reference compares the physical and the right index of every tensor; kept as is)."""
        return max(max(int(t.shape[0]), int(t.shape[2])) for t in self.tensors)

    def get_total_bond(self) -> int:
        """Sum of the internal bonds (mps.py:567-578)."""
        return sum(int(t.shape[1]) for t in self.tensors[1:])

    def get_cost(self) -> int:
        """Sum of chi^3 over the internal bonds (mps.py:580-591)."""
        return sum(int(t.shape[1]) ** 3 for t in self.tensors[1:])

    def check_if_valid_mps(self) -> None:
        """Neighbouring bonds must agree (mps.py:1567-1579)."""
        for a, b in zip(self.tensors[:-1], self.tensors[1:]):
            assert a.shape[2] == b.shape[1], "bond dimensions of neighbouring tensors differ"

    def almost_equal(self, other: "MPS") -> bool:
        """Tensor-by-tensor comparison, not a comparison of the represented vectors (mps.py:700-717)."""
        return (self.length == other.length
                and all(a.shape == b.shape and np.allclose(a, b) for a, b in zip(self.tensors, other.tensors)))

    def to_vec(self) -> np.ndarray:
        """Dense state vector in the reference's index order (mps.py:1633-1660: site 0 is the fastest-running index)."""
        acc = np.ones((1, 1), dtype=C)                                  # (physical multi-index, open left bond), built from the right end
        for t in reversed(self.tensors):
            acc = np.einsum("slr,xr->xsl", t, acc).reshape(-1, t.shape[1])
        return acc[:, 0].copy()

    def _cut_values(self, sites) -> np.ndarray | None:
        assert len(sites) == 2 and sites[0] + 1 == sites[1], "defined on a bond (two adjacent sites)"
        a, b = self.tensors[sites[0]], self.tensors[sites[1]]
        if a.shape[2] == 1:
            return None
        theta = np.einsum("slk,tkr->lstr", a, b).reshape(a.shape[1] * a.shape[0], b.shape[0] * b.shape[2])
        return np.linalg.svd(theta, compute_uv=False)

    def get_entropy(self, sites) -> np.float64:
        """Von Neumann entropy of the two-site block's singular values (mps.py:604-642)."""
        sv = self._cut_values(list(sites))
        if sv is None:
            return np.float64(0.0)
        w = sv.astype(np.float64) ** 2
        total = np.sum(w, dtype=np.float64)
        if total == 0.0:
            return np.float64(0.0)
        prob = w / total
        return np.float64(-np.sum(prob * np.log(prob + np.finfo(np.float64).tiny), dtype=np.float64))

    def get_schmidt_spectrum(self, sites) -> np.ndarray:
        """The block's singular values in a NaN-padded vector of 500 (mps.py:644-678)."""
        out = np.full(500, np.nan)
        sv = self._cut_values(list(sites))
        if sv is None:
            out[0] = 1.0
        else:
            out[: min(500, len(sv))] = sv[:500]
        return out

    def scalar_product(self, other: "MPS", sites=None) -> np.complex128:
        """<self|other>: over the whole chain, or the local contraction at one site / one adjacent pair (mps.py:901-959)."""
        if sites is None:
            env = np.ones((1, 1), dtype=C)                               # (bra bond, ket bond)
            for a, b in zip(self.tensors, other.tensors):
                env = np.einsum("xy,sxp,syq->pq", env, a.conj(), b)
            return np.complex128(env[0, 0])
        where = [sites] if isinstance(sites, (int, np.integer)) else list(sites)
        if len(where) == 1:
            return np.complex128(np.vdot(self.tensors[where[0]], other.tensors[where[0]]))
        if len(where) == 2:
            i, j = where
            assert j == i + 1, "Only nearest-neighbor two-site overlaps supported."
            bra = np.einsum("slk,tkr->sltr", self.tensors[i], self.tensors[j])
            ket = np.einsum("slk,tkr->sltr", other.tensors[i], other.tensors[j])
            return np.complex128(np.vdot(bra, ket))
        raise ValueError(f"Invalid `sites` argument: {sites!r}")

    def norm(self, site: int | None = None) -> np.float64:
        """<psi|psi> (the reference returns the squared norm, mps.py:1539-1565); with ``site`` the local contraction is taken where the
        tracked centre sits on that site, which is the same number for a canonical state."""
        if site is not None and self.orthogonality_center is not None and site == self.orthogonality_center:
            return np.float64(self.scalar_product(self, site).real)
        return np.float64(self.scalar_product(self).real)

    def _operator_factors(self, observable):
        """(sites ascending, per-site operator factors [k][2][2]): a one-site matrix as is, a two-site matrix split into a sum of
        products (operator Schmidt decomposition), so that any two sites -- adjacent or not -- contract along the chain."""
        raw = observable.sites
        where = [int(raw)] if isinstance(raw, (int, np.integer)) else [int(v) for v in raw]
        assert len(where) < 3, "Only one- and two-site observables are currently implemented."
        for site in where:
            assert 0 <= site < self.length, f"Observable acting on non-existing site: {site}"
        m = np.asarray(observable.gate.matrix, dtype=C)
        if len(where) == 1:
            if m.shape != (2, 2):
                raise ValueError(f"Local observable matrix shape {m.shape} does not match site {where[0]} dimension 2.")
            return where, [m[None]]
        if m.shape != (4, 4):
            raise ValueError(f"Two-site observable matrix shape {m.shape} does not match site dimensions 2 and 2.")
        t = m.reshape(2, 2, 2, 2)                                        # (out_0, out_1, in_0, in_1) in ``sites`` order
        if where[0] > where[1]:
            where, t = where[::-1], t.transpose(1, 0, 3, 2)
        u, sv, vh = np.linalg.svd(t.transpose(0, 2, 1, 3).reshape(4, 4))
        return where, [(u * sv).T.reshape(4, 2, 2), vh.reshape(4, 2, 2)]

    def expect(self, observable) -> np.float64:
        """Re <psi|O|psi> for a one- or two-site observable (mps.py:1236-1280; not divided by the norm, like the reference)."""
        where, factors = self._operator_factors(observable)
        nk = factors[0].shape[0]
        env = np.zeros((nk, 1, 1), dtype=C)                              # one transfer chain per term of the operator sum
        env[:, 0, 0] = 1.0
        for site, t in enumerate(self.tensors):
            if site in where:
                op = factors[where.index(site)]
                env = np.einsum("kxy,sxp,kst,tyq->kpq", env, t.conj(), op, t)
            else:
                env = np.einsum("kxy,sxp,syq->kpq", env, t.conj(), t)
        val = complex(env[:, 0, 0].sum())
        assert val.imag < 1e-13, f"Measurement should be real, '{val.real:16f}+{val.imag:16f}i'."
        return np.float64(val.real)

    def local_expect(self, operator, sites) -> np.complex128:
        """The contraction restricted to the observable's site(s) -- the expectation value when the centre covers them
        (mps.py:961-1047); one site or an adjacent pair."""
        where = [sites] if isinstance(sites, (int, np.integer)) else list(sites)
        m = np.asarray(operator.gate.matrix, dtype=C)
        if len(where) == 1:
            a = self.tensors[where[0]]
            return np.complex128(np.vdot(a, np.einsum("st,tlr->slr", m, a)))
        if len(where) == 2 and where[1] == where[0] + 1:
            theta = np.einsum("slk,tkr->stlr", self.tensors[where[0]], self.tensors[where[1]])
            acted = np.einsum("abst,stlr->ablr", m.reshape(2, 2, 2, 2), theta)
            return np.complex128(np.vdot(theta, acted))
        raise ValueError("Local observable must be one-site or nearest-neighbor two-site.")

    # ------------------------------------------------------------------------------------------------------------------------------
    # Gauge moves, compression and projective sampling of a host MPS (mps.py:355-407, 719-899, 1282-1493, 1581-1631): numpy, written
    # around ONE primitive -- push the non-isometric factor of a site into its neighbour by a QR -- so that every move leaves the
    # represented vector untouched; only ``compress`` truncates.
    # ------------------------------------------------------------------------------------------------------------------------------
    def _push(self, site: int, to_right: bool) -> None:
        t = self.tensors[site]
        d, l, r = t.shape
        if to_right:
            q, rest = np.linalg.qr(t.reshape(d * l, r))                    # rows (s, l): the site becomes a left isometry
            self.tensors[site] = q.reshape(d, l, -1)
            self.tensors[site + 1] = np.einsum("kr,srq->skq", rest, self.tensors[site + 1])
        else:
            q, rest = np.linalg.qr(t.transpose(0, 2, 1).reshape(d * r, l))
            self.tensors[site] = q.reshape(d, r, -1).transpose(0, 2, 1)
            self.tensors[site - 1] = np.einsum("slk,qk->slq", self.tensors[site - 1], rest)

    def set_canonical_form(self, orthogonality_center: int, decomposition: str = "QR") -> None:
        """Mixed-canonical form around ``orthogonality_center`` whatever the current gauge (mps.py:766-813); both decompositions are
        carried out without truncation here."""
        if not 0 <= orthogonality_center < self.length:
            raise ValueError(f"orthogonality_center {orthogonality_center} outside the chain of length {self.length}.")
        for site in range(orthogonality_center):
            self._push(site, to_right=True)
        for site in range(self.length - 1, orthogonality_center, -1):
            self._push(site, to_right=False)
        self.orthogonality_center = orthogonality_center

    def set_center(self, site: int | None) -> None:
        """Declare where the centre is (mps.py:303-314): bookkeeping only."""
        self.orthogonality_center = site

    def shift_center_to(self, site: int, decomposition: str = "QR") -> None:
        """Move the tracked centre (mps.py:355-372)."""
        if self.orthogonality_center is None:
            raise ValueError("shift_center_to: the gauge is unknown; call set_canonical_form first.")
        while self.orthogonality_center < site:
            self._push(self.orthogonality_center, to_right=True)
            self.orthogonality_center += 1
        while self.orthogonality_center > site:
            self._push(self.orthogonality_center, to_right=False)
            self.orthogonality_center -= 1

    def flip_network(self) -> None:
        """Reverse the chain (mps.py:680-698): site order and the two bond indices of every tensor."""
        self.tensors = [t.transpose(0, 2, 1).copy() for t in reversed(self.tensors)]
        self.physical_dimensions = list(reversed(self.physical_dimensions))
        self.flipped = not self.flipped
        if self.orthogonality_center is not None:
            self.orthogonality_center = self.length - 1 - self.orthogonality_center

    def check_canonical_form(self) -> list[int]:
        """The sites that can serve as orthogonality centre: everything to their left a left isometry, everything to their right a
        right isometry, tested with numpy's allclose like the reference (mps.py:1581-1631)."""
        left = [np.allclose(np.einsum("slk,slq->kq", t.conj(), t), np.eye(t.shape[2])) for t in self.tensors]
        right = [np.allclose(np.einsum("slk,sqk->lq", t, t.conj()), np.eye(t.shape[1])) for t in self.tensors]
        return [i for i in range(self.length) if all(left[:i]) and all(right[i + 1:])]

    @staticmethod
    def _keep(sv: np.ndarray, mode: str, threshold: float, max_bond_dim: int | None, min_keep: int = 1) -> int:
        """How many singular values a split keeps (core/linalg/svd_utils.py:22-104; the device version is csrc/yb_truncate.cuh)."""
        n = int(sv.size)
        if n == 0:
            return 0
        if mode == "hard_cutoff":
            keep = int(np.sum(sv > threshold))
        elif mode == "relative":
            keep = 0 if sv[0] <= 0.0 else int(np.sum(sv / sv[0] >= threshold))
        elif mode in ("discarded_weight", "relative_discarded_weight"):
            w = np.asarray(sv, dtype=np.float64) ** 2
            if mode == "relative_discarded_weight":
                if sv[0] <= 0.0:
                    w, keep = None, 0
                else:
                    w = (np.asarray(sv, dtype=np.float64) / float(sv[0])) ** 2
                    w = w / float(np.sum(w))
            if w is not None:
                tail = np.cumsum(w[::-1])                                # weight of the k smallest, k = 1 .. n
                if mode == "discarded_weight":
                    hit = np.nonzero(tail >= threshold)[0]
                    keep = n if hit.size == 0 else max(n - int(hit[0]), min_keep)
                else:
                    keep = n - int(np.sum(tail <= threshold))
        else:
            raise ValueError(f"Unknown truncation mode: {mode!r}")
        if max_bond_dim is not None:
            keep = min(keep, max_bond_dim)
        return min(max(keep, min_keep), n)

    def compress(self, threshold: float, *, max_bond_dim: int | None = None, trunc_mode: str = "discarded_weight") -> None:
        """Right-canonicalise without truncation, truncate bond by bond from the left with the singular values to the right, put the
        centre back where it was (mps.py:841-899)."""
        if self.length == 1:
            return
        if self.orthogonality_center is not None:
            back = self.orthogonality_center
        else:
            sites = self.check_canonical_form()
            back = sites[0] if sites else self.length // 2
        self.set_canonical_form(0)
        for site in range(self.length - 1):
            a, b = self.tensors[site], self.tensors[site + 1]
            theta = np.einsum("slk,tkr->sltr", a, b).reshape(a.shape[0] * a.shape[1], b.shape[0] * b.shape[2])
            u, sv, vh = np.linalg.svd(theta, full_matrices=False)
            keep = self._keep(sv, trunc_mode, threshold, max_bond_dim)
            self.tensors[site] = u[:, :keep].reshape(a.shape[0], a.shape[1], keep)
            self.tensors[site + 1] = (sv[:keep, None] * vh[:keep]).reshape(keep, b.shape[0], b.shape[2]).transpose(1, 0, 2)
            self.orthogonality_center = site + 1
        self.shift_center_to(back)

    _BASIS_ROTATIONS = {"Z": np.eye(2, dtype=C), "X": np.array([[1, 1], [1, -1]], dtype=C) / np.sqrt(2),
                        "Y": np.array([[1, -1j], [1, 1j]], dtype=C) / np.sqrt(2)}

    def measure_single_shot(self, basis: str = "Z", rng: np.random.Generator | None = None) -> int:
        """One projective measurement of every site, left to right, on a copy (mps.py:1282-1341): the outcome of site k conditions the
        tensor of site k + 1; bit k of the returned integer is the outcome of site k.  One ``rng.choice`` per site, like the reference."""
        rot = self._BASIS_ROTATIONS.get(basis.upper())
        if rot is None:
            raise ValueError(f"Invalid basis: {basis}. Expected 'X', 'Y', or 'Z'.")
        rng = np.random.default_rng() if rng is None else rng
        work = MPS.__new__(MPS)
        work.length, work.physical_dimensions, work.flipped = self.length, list(self.physical_dimensions), self.flipped
        work.tensors, work.orthogonality_center = [t.copy() for t in self.tensors], self.orthogonality_center
        if work.orthogonality_center is None:
            work.set_canonical_form(0)
        outcome = 0
        for site in range(work.length):
            work.shift_center_to(site)
            rotated = np.einsum("ab,blr->alr", rot, work.tensors[site])
            prob = np.einsum("alr,alr->a", rotated, rotated.conj()).real.copy()
            prob /= np.sum(prob)
            bit = int(rng.choice(len(prob), p=prob))
            outcome |= bit << site
            if site != work.length - 1:
                work.tensors[site + 1] = np.einsum("lk,skr->slr", rotated[bit], work.tensors[site + 1]) / np.sqrt(prob[bit])
                work.orthogonality_center = site + 1
        return outcome

    def measure_shots(self, shots: int, basis: str = "Z", rng: np.random.Generator | None = None) -> dict[int, int]:
        """Histogram of ``shots`` single shots (mps.py:1343-1417; the reference fans the shots out over a process pool with unseeded
        generators -- here they are taken one after the other, from ``rng`` when one is given)."""
        counts: dict[int, int] = {}
        for _ in range(max(int(shots), 1)):
            outcome = self.measure_single_shot(basis, rng)
            counts[outcome] = counts.get(outcome, 0) + 1
        return counts

    def project_onto_bitstring(self, bitstring: str) -> np.complex128:
        """|<bitstring|psi>|^2, site 0 first (mps.py:1495-1537)."""
        assert len(bitstring) == self.length, "Bitstring length must match number of sites"
        row = np.ones(1, dtype=C)
        for site, char in enumerate(bitstring):
            index = int(char)
            assert 0 <= index < 2, f"Invalid state index {index} at site {site}"
            row = row @ self.tensors[site][index]
        return np.complex128(abs(row[0]) ** 2)


class MPO:
    """Matrix product operator: tensors (s_out, s_in, l, r).  Factories build the operator with an exact finite-state
    construction (bond dimension 2 + number of two-body terms); the reference reaches the same operator through an FSM +
    SVD compression (mpo.py:1171-1423), i.e. a different MPO gauge with identical H."""

    def __init__(self) -> None:
        self.tensors: list[np.ndarray] = []
        self.length = 0
        self.physical_dimension = 2

    @classmethod
    def custom(cls, tensors) -> "MPO":
        m = cls()
        m.tensors = [np.asarray(t, dtype=C) for t in tensors]
        m.length = len(m.tensors)
        for t in m.tensors:
            if t.ndim != 4 or t.shape[0] != 2 or t.shape[1] != 2:
                raise ValueError("MPO tensors must have shape (2, 2, Dl, Dr)")
        return m

    @classmethod
    def pauli(cls, *, length: int, two_body=None, one_body=None, bc: str = "open", **_unused) -> "MPO":
        """H = sum_i sum_k c_k A_k(i) B_k(i+1) + sum_i sum_m c_m P_m(i)   (mpo.py:248-325)."""
        if bc not in ("open", "periodic"):
            raise ValueError("bc must be 'open' or 'periodic'")
        two = [(complex(c), _PAULI[a.upper()], _PAULI[b.upper()]) for c, a, b in (two_body or []) if c != 0]
        one = sum((complex(c) * _PAULI[p.upper()] for c, p in (one_body or [])), np.zeros((2, 2), dtype=C))
        nt = len(two)
        periodic = bc == "periodic" and length > 2 and nt > 0
        dim = 2 + nt + (nt if periodic else 0)
        last = dim - 1
        w = np.zeros((dim, dim, 2, 2), dtype=C)
        w[0, 0] = _PAULI["I"]
        w[last, last] = _PAULI["I"]
        w[last, 0] = one
        for k, (c, a, b) in enumerate(two):
            w[1 + k, 0] = b
            w[last, 1 + k] = c * a
            if periodic:                       # channel carrying B_k(0) through the chain to c_k A_k(L-1)
                w[1 + nt + k, 1 + nt + k] = _PAULI["I"]
        tensors = []
        for i in range(length):
            if length == 1:
                t = w[last:last + 1, 0:1].copy()
            elif i == 0:
                t = w[last:last + 1, :].copy()
                if periodic:
                    for k, (c, a, b) in enumerate(two):
                        t[0, 1 + nt + k] = b
            elif i == length - 1:
                t = w[:, 0:1].copy()
                if periodic:
                    for k, (c, a, b) in enumerate(two):
                        t[1 + nt + k, 0] = c * a
            else:
                t = w
            tensors.append(np.ascontiguousarray(t.transpose(2, 3, 0, 1)))
        return cls.custom(tensors)

    _TOKEN = re.compile(r"([IXYZixyz])\s*(\d+)")

    @classmethod
    def _parse_pauli_string(cls, spec: str) -> dict[int, str]:
        """"X0 Y2 Z5" (blank- or comma-separated) -> {site: label}; "" is the identity (mpo.py:2093-2133)."""
        text = spec.replace(",", " ").strip()
        if not text:
            return {}
        out: dict[int, str] = {}
        for op, idx in cls._TOKEN.findall(text):
            site = int(idx)
            if site in out:
                raise ValueError(f"Duplicate site {site} in spec '{spec}'.")
            out[site] = op.upper()
        if cls._TOKEN.sub("", text).split():
            raise ValueError(f"Invalid token(s) in spec '{spec}'. Use forms like 'X0 Y2 Z5'.")
        return out

    def from_pauli_sum(self, *, terms, length: int, physical_dimension: int = 2, tol: float = 1e-12, max_bond_dim: int | None = None,
                       n_sweeps: int = 2) -> None:
        """H = sum_t c_t P_t for Pauli strings of any range (mpo.py:1171-1317), filled into this MPO.

        Exact finite-state construction, no compression: one "idle" state (identities until a term starts), one "done" state
        (identities after it ended), and between the first and the last non-identity site of a term one state per distinct REMAINING
        suffix -- terms that still have the same operators to apply share a channel, whatever came before (the coefficient and the
        first operator sit on the transition out of "idle").  Bond dimension at a cut = 2 + number of distinct suffixes crossing it:
        3 for the Ising chain, L - b + 1 for all-to-all Z_i Z_j.  The reference builds a suffix automaton over whole strings and
        compresses it with SVD sweeps (``tol``, ``max_bond_dim``, ``n_sweeps`` are accepted for its call signature; the exact
        automaton needs none of them): same operator, different MPO gauge -- H_eff does not see the gauge."""
        if physical_dimension != 2:
            raise ValueError("Only physical_dimension=2 is supported by this Pauli MPO builder.")
        if length <= 0:
            raise ValueError("length must be positive.")
        del tol, max_bond_dim, n_sweeps
        self.length, self.physical_dimension = int(length), 2
        parsed = []
        for coeff, spec in terms:
            ops = self._parse_pauli_string(spec)
            for site in ops:
                if not 0 <= site < length:
                    raise ValueError(f"Site index {site} outside [0, {length - 1}].")
            support = tuple(sorted((site, lab) for site, lab in ops.items() if lab != "I"))
            parsed.append((complex(coeff), support))
        # states per internal bond b (between sites b - 1 and b): 0 = done, 1 = idle, 2... = remaining suffixes crossing the bond
        suffix_ids: list[dict] = [dict() for _ in range(length + 1)]
        for _, support in parsed:
            for k in range(1, len(support)):
                rest = support[k:]
                for b in range(support[k - 1][0] + 1, support[k][0] + 1):
                    suffix_ids[b].setdefault(rest, 2 + len(suffix_ids[b]))
        dims = [1] + [2 + len(suffix_ids[b]) for b in range(1, length)] + [1]
        DONE, IDLE = 0, 1

        def left_index(b: int, state) -> int:               # index of a state on bond b (the chain ends have one state each)
            if b == 0 or b == length:
                return 0
            return state if isinstance(state, int) else suffix_ids[b][state]

        tensors = [np.zeros((2, 2, dims[i], dims[i + 1]), dtype=C) for i in range(length)]
        eye = _PAULI["I"]
        for i in range(length):
            w = tensors[i]
            if i > 0 and i < length - 1:
                w[:, :, DONE, DONE] += eye
                w[:, :, IDLE, IDLE] += eye
            elif length > 1 and i == 0:
                w[:, :, 0, IDLE] += eye                      # nothing has started yet
            elif length > 1:
                w[:, :, DONE, 0] += eye                      # everything has ended
            for rest, idx in suffix_ids[i].items():          # channels entering site i from the left
                site, lab = rest[0]
                if site > i:
                    w[:, :, idx, left_index(i + 1, rest)] += eye
                else:
                    nxt = rest[1:]
                    w[:, :, idx, left_index(i + 1, nxt if nxt else DONE)] += _PAULI[lab]
        for coeff, support in parsed:                        # transitions out of "idle": coefficient times the first operator
            if not support:
                site, mat, nxt = 0, coeff * eye, ()
            else:
                site, mat, nxt = support[0][0], coeff * _PAULI[support[0][1]], support[1:]
            if not support and length > 1:
                tensors[0][:, :, 0, DONE] += mat             # a constant: "done" right away, identities from there on
                continue
            tensors[site][:, :, left_index(site, IDLE), left_index(site + 1, nxt if nxt else DONE)] += mat
        self.tensors = tensors

    @classmethod
    def from_matrix(cls, mat, d: int = 2, max_bond: int | None = None, cutoff: float = 1e-12) -> "MPO":
        """Factorise a dense operator on n qubits into an MPO by successive SVDs, site 0 first (mpo.py:1852-1974): singular values
        <= ``cutoff`` are dropped, at most ``max_bond`` kept."""
        if d != 2:
            raise NotImplementedError("the B200 TJM path supports qubits (d = 2) only")
        m = np.asarray(mat, dtype=C)
        if m.ndim != 2 or m.shape[0] != m.shape[1]:
            raise ValueError("mat must be a square 2-D array.")
        n = int(round(np.log2(m.shape[0])))
        if 2 ** n != m.shape[0] or n < 1:
            raise ValueError(f"matrix dimension {m.shape[0]} is not a power of d = 2.")
        # (out_0 .. out_{n-1}, in_0 .. in_{n-1}) -> (out_0, in_0, out_1, in_1, ...): one (out, in) pair per site
        t = m.reshape((2,) * (2 * n)).transpose([v for k in range(n) for v in (k, n + k)])
        tensors, left = [], 1
        rest = t.reshape(left * 4, -1)
        for _site in range(n - 1):
            u, sv, vh = np.linalg.svd(rest, full_matrices=False)
            keep = max(1, int(np.sum(sv > cutoff)))
            if max_bond is not None:
                keep = min(keep, int(max_bond))
            tensors.append(np.ascontiguousarray(u[:, :keep].reshape(left, 2, 2, keep).transpose(1, 2, 0, 3)))
            rest = (sv[:keep, None] * vh[:keep]).reshape(keep * 4, -1)
            left = keep
        tensors.append(np.ascontiguousarray(rest.reshape(left, 2, 2, 1).transpose(1, 2, 0, 3)))
        return cls.custom(tensors)

    @classmethod
    def ising(cls, length: int, J: float, g: float, *, bc: str = "open", **kw) -> "MPO":   # noqa: N803
        """-J sum Z Z - g sum X   (mpo.py:327-364)."""
        return cls.pauli(length=length, two_body=[(-J, "Z", "Z")], one_body=[(-g, "X")], bc=bc)

    @classmethod
    def heisenberg(cls, length: int, Jx: float, Jy: float, Jz: float, h: float = 0.0, *, bc: str = "open", **kw) -> "MPO":   # noqa: N803
        """-(Jx XX + Jy YY + Jz ZZ) - h sum Z   (mpo.py:366-404)."""
        return cls.pauli(length=length, two_body=[(-Jx, "X", "X"), (-Jy, "Y", "Y"), (-Jz, "Z", "Z")],
                         one_body=[(-h, "Z")] if h != 0 else [], bc=bc)

    @classmethod
    def fermi_hubbard_1d(cls, length: int, t: float, u: float, *, jordan_wigner: bool = False) -> "MPO":
        """1-D Fermi-Hubbard chain as a qubit Hamiltonian (mpo.py:409-470, ``jordan_wigner=True``): spin orbitals interleaved 1 up, 1 down,
        2 up, 2 down, ... (``length`` = number of spin orbitals, even);  U n_up n_down = U/4 (I - Z_up - Z_down + Z_up Z_down) on every
        site and hops -t/2 (X Z X + Y Z Y) between equal spins of neighbouring sites, the Z on the orbital in between.  The fermionic
        form on four-level sites is outside the qubit path."""
        if not jordan_wigner:
            raise NotImplementedError("the B200 TJM path supports qubits (d = 2) only: pass jordan_wigner=True")
        if length % 2 != 0 or length < 2:
            raise ValueError("length must be an even integer \u2265 2 (ordering: 1\u2191,1\u2193,2\u2191,2\u2193,...).")
        terms = []
        for up in range(0, length, 2):
            terms += [(u / 4, ""), (-u / 4, f"Z{up}"), (-u / 4, f"Z{up + 1}"), (u / 4, f"Z{up} Z{up + 1}")]
        for a in range(length - 2):
            terms += [(-t / 2, f"{p}{a} Z{a + 1} {p}{a + 2}") for p in "XY"]
        mpo = cls()
        mpo.from_pauli_sum(terms=terms, length=length)
        return mpo

    @classmethod
    def identity(cls, length: int, physical_dimension: int = 2) -> "MPO":
        """mpo.py:1015-1028."""
        if physical_dimension != 2:
            raise NotImplementedError("the B200 TJM path supports qubits (d = 2) only")
        return cls.custom([np.eye(2, dtype=C).reshape(2, 2, 1, 1) for _ in range(length)])

    @classmethod
    def mpo_sum(cls, mpos) -> "MPO":
        """Sum of MPOs of equal length as the direct sum of their bond spaces (mpo.py:1976-2058)."""
        mpos = list(mpos)
        if not mpos:
            raise ValueError("mpo_sum needs at least one MPO.")
        n = mpos[0].length
        if any(m.length != n for m in mpos):
            raise ValueError("MPO lengths must match.")
        out = []
        for i in range(n):
            blocks = [m.tensors[i] for m in mpos]
            if n == 1:
                out.append(sum(blocks))
                continue
            dl = 1 if i == 0 else sum(b.shape[2] for b in blocks)
            dr = 1 if i == n - 1 else sum(b.shape[3] for b in blocks)
            w = np.zeros((2, 2, dl, dr), dtype=C)
            lo, ro = 0, 0
            for b in blocks:
                ls = slice(0, 1) if i == 0 else slice(lo, lo + b.shape[2])
                rs = slice(0, 1) if i == n - 1 else slice(ro, ro + b.shape[3])
                w[:, :, ls, rs] += b
                lo, ro = lo + b.shape[2], ro + b.shape[3]
            out.append(w)
        return cls.custom(out)

    def __add__(self, other: "MPO") -> "MPO":
        return MPO.mpo_sum([self, other])

    def check_if_valid_mpo(self) -> bool:
        """Neighbouring bonds agree and the chain ends are closed (mpo.py:2060-2076)."""
        ok = all(a.shape[3] == b.shape[2] for a, b in zip(self.tensors[:-1], self.tensors[1:]))
        return bool(ok and self.tensors and self.tensors[0].shape[2] == 1 and self.tensors[-1].shape[3] == 1)

    def to_sparse_matrix(self):
        import scipy.sparse
        return scipy.sparse.csr_matrix(self.to_matrix())

    def to_matrix(self) -> np.ndarray:
        acc = self.tensors[0][:, :, 0, :]
        for w in self.tensors[1:]:
            acc = np.einsum("oir,pqrs->opiqs", acc, w)
            s = acc.shape
            acc = acc.reshape(s[0] * s[1], s[2] * s[3], s[4])
        return acc[:, :, 0]


class State:
    """Initial state specification for ``Simulator.run`` (state.py:35-60); only ``representation="mps"`` is in scope."""

    def __init__(self, length: int | None = None, *, initial: str = "zeros", representation: str | None = None,
                 physical_dimensions=None, tensors=None, vector=None, density_matrix=None, pad: int | None = None,
                 basis_string: str | None = None, seed: int | None = None) -> None:
        if length is not None and length <= 0:
            raise ValueError("length must be a positive integer.")
        if sum(x is not None for x in (tensors, vector, density_matrix)) > 1:
            raise ValueError("Specify at most one of tensors, vector, and density_matrix.")
        if vector is not None or density_matrix is not None or (representation not in (None, "mps")):
            raise NotImplementedError("the B200 TJM path takes representation='mps' states only (vector / density-matrix "
                                      "backends are outside SURVEY.md section 8)")
        self.representation = "mps"
        self.initial, self.pad, self.basis_string, self.seed = initial, pad, basis_string, seed
        if tensors is not None:
            if len(tensors) == 0:
                raise ValueError("tensors must be a non-empty list of MPS cores.")
            if length is not None and length != len(tensors):
                raise ValueError(f"length={length} does not match len(tensors)={len(tensors)}.")
            self.length = len(tensors)
            self.mps = MPS(self.length, tensors=tensors)
        else:
            if length is None:
                raise ValueError("length is required for preset states.")
            self.length = length
            self.mps = MPS(length, state=initial, pad=pad, basis_string=basis_string, physical_dimensions=physical_dimensions)
        self.physical_dimensions = self.mps.physical_dimensions

    @classmethod
    def from_mps(cls, mps: MPS) -> "State":
        """Wrap an existing MPS (state.py:175-188): the State references ``mps`` itself."""
        wrapped = cls.__new__(cls)
        wrapped.representation, wrapped.initial, wrapped.pad, wrapped.basis_string, wrapped.seed = "mps", "zeros", None, None, None
        wrapped.length, wrapped.mps, wrapped.physical_dimensions = mps.length, mps, mps.physical_dimensions
        return wrapped

    def ensure_encoded(self, representation: str = "mps") -> "State":
        if representation != "mps":
            raise NotImplementedError("the B200 TJM path takes representation='mps' states only")
        return self


class Hamiltonian:
    """Analog operator wrapper (hamiltonian.py:36-441): factories delegate to :class:`MPO`; ``piecewise`` switches static
    pieces on the analog ``dt`` grid (:176-263)."""

    def __init__(self, length=None, *, tensors=None, matrix=None, sparse_matrix=None, physical_dimension: int = 2) -> None:
        """The reference's manual constructor (hamiltonian.py:49-167): exactly one of ``tensors`` (MPO cores in the reference's
        (left, right, out, in) index order, ``MPO.custom`` :1146-1168), ``matrix`` (dense, factorised by ``MPO.from_matrix``) or
        ``sparse_matrix`` (densified, then the same).  The factories pass a ready MPO as the first argument."""
        self._pieces: tuple | None = None
        if isinstance(length, MPO):
            if tensors is not None or matrix is not None or sparse_matrix is not None:
                raise ValueError("Pass exactly one of tensors, matrix, or sparse_matrix, or use a classmethod preset.")
            self._mpo, self.length = length, length.length
            return
        if physical_dimension != 2:
            if physical_dimension <= 0:
                raise ValueError("physical_dimension must be a positive integer.")
            raise NotImplementedError("the B200 TJM path supports qubits (d = 2) only")
        if sum(x is not None for x in (tensors, matrix, sparse_matrix)) != 1:
            raise ValueError("Pass exactly one of tensors, matrix, or sparse_matrix, or use a classmethod preset.")
        if tensors is not None:
            if len(tensors) == 0:
                raise ValueError("tensors must be a non-empty list of MPO cores.")
            if length is not None and length != len(tensors):
                raise ValueError(f"length={length} does not match len(tensors)={len(tensors)}.")
            self._mpo = MPO.custom([np.transpose(np.asarray(t, dtype=C), (2, 3, 0, 1)) for t in tensors])
        else:
            dense = matrix if matrix is not None else (sparse_matrix.toarray() if hasattr(sparse_matrix, "toarray") else np.asarray(sparse_matrix))
            dense = np.asarray(dense, dtype=C)
            label = "matrix" if matrix is not None else "sparse_matrix"
            if dense.ndim != 2 or dense.shape[0] != dense.shape[1]:
                raise ValueError("matrix must be a square 2-D array." if matrix is not None else "sparse_matrix must be square.")
            if length is not None and dense.shape[0] != 2 ** length:
                raise ValueError(f"{label} dimension {dense.shape[0]} does not match physical_dimension**length={2 ** length}.")
            self._mpo = MPO.from_matrix(dense, 2)
        self.length = self._mpo.length

    @classmethod
    def from_mpo(cls, mpo: MPO) -> "Hamiltonian":
        return cls(mpo)

    @classmethod
    def ising(cls, length: int, J: float, g: float, **kw) -> "Hamiltonian":   # noqa: N803
        return cls(MPO.ising(length, J, g, **kw))

    @classmethod
    def heisenberg(cls, length: int, Jx: float, Jy: float, Jz: float, h: float = 0.0, **kw) -> "Hamiltonian":   # noqa: N803
        return cls(MPO.heisenberg(length, Jx, Jy, Jz, h, **kw))

    @classmethod
    def pauli(cls, **kw) -> "Hamiltonian":
        return cls(MPO.pauli(**kw))

    @classmethod
    def fermi_hubbard_1d(cls, length: int, t: float, u: float, *, jordan_wigner: bool = False) -> "Hamiltonian":
        return cls(MPO.fermi_hubbard_1d(length, t, u, jordan_wigner=jordan_wigner))

    @classmethod
    def piecewise(cls, pieces) -> "Hamiltonian":
        """hamiltonian.py:176-232: ``[(Hamiltonian, duration), ...]``; durations must be positive multiples of ``dt`` that sum to
        ``elapsed_time`` (checked when the run expands them to one MPO per interval)."""
        from collections.abc import Sequence as _Seq
        if isinstance(pieces, (str, bytes)) or not isinstance(pieces, _Seq):
            raise TypeError("pieces must be a non-empty sequence of (Hamiltonian, duration) pairs.")
        raw = tuple(pieces)
        if not raw:
            raise ValueError("pieces must be a non-empty sequence of (Hamiltonian, duration) pairs.")
        normalized, length = [], None
        for index, item in enumerate(raw):
            if not isinstance(item, tuple) or len(item) != 2:
                raise TypeError(f"pieces[{index}] must be a (Hamiltonian, duration) tuple.")
            ham, duration = item
            if not isinstance(ham, Hamiltonian):
                raise TypeError(f"pieces[{index}] must start with a Hamiltonian, got {type(ham).__name__}.")
            if ham.is_piecewise:
                raise ValueError("pieces must be static Hamiltonians, not nested piecewise Hamiltonians.")
            if isinstance(duration, bool) or not isinstance(duration, (int, float, np.floating, np.integer)):
                raise TypeError(f"pieces[{index}] duration must be a real number, got {type(duration).__name__}.")
            dur = float(duration)
            if not np.isfinite(dur) or dur <= 0.0:
                raise ValueError(f"pieces[{index}] duration must be finite and positive, got {dur}.")
            if length is None:
                length = ham.length
            elif ham.length != length:
                raise ValueError(f"pieces[{index}] length {ham.length} does not match piece 0 length {length}.")
            normalized.append((ham, dur))
        wrapped = cls.__new__(cls)
        wrapped._mpo, wrapped._pieces, wrapped.length = None, tuple(normalized), length
        return wrapped

    @property
    def is_piecewise(self) -> bool:
        return getattr(self, "_pieces", None) is not None

    @property
    def pieces(self) -> tuple:
        if self._pieces is None:
            raise ValueError("Static Hamiltonians do not have piecewise durations.")
        return self._pieces

    @property
    def duration(self) -> float:
        if self._pieces is None:
            raise ValueError("Static Hamiltonians do not have a piecewise duration.")
        return float(sum(d for _, d in self._pieces))

    def ensure_mpo(self) -> "Hamiltonian":
        if self.is_piecewise:
            raise ValueError("A piecewise Hamiltonian has no single static operator.")
        return self

    @property
    def mpo(self) -> MPO:
        if self.is_piecewise:
            raise ValueError("A piecewise Hamiltonian has no single static operator.")
        return self._mpo

    def to_matrix(self) -> np.ndarray:
        return self.mpo.to_matrix()

    @property
    def matrix(self) -> np.ndarray:
        return self.to_matrix()

    def to_sparse_matrix(self):
        return self.mpo.to_sparse_matrix()

    @property
    def sparse_matrix(self):
        return self.to_sparse_matrix()

    def ensure_sparse(self) -> "Hamiltonian":
        return self.ensure_mpo()


@dataclass
class Result:
    """Outputs of ``Simulator.run`` (result.py:155-189): rows follow the user's observable order."""
    sim_params: object = None
    observables: list = field(default_factory=list)
    expectation_values: list = field(default_factory=list)
    trajectories: list = field(default_factory=list)
    times: np.ndarray | None = None
    runtime_cost: np.ndarray | None = None
    max_bond: np.ndarray | None = None
    total_bond: np.ndarray | None = None
    noise_model: object = None
    output_state: object = None
    multi_time_times: np.ndarray | None = None     # (result.py:182-183; filled by the unitary-ensemble path only, which is out of scope here)
    multi_time_results: np.ndarray | None = None
    measurements: list = field(default_factory=list)
    counts: dict | None = None
    status: np.ndarray | None = None          # per-trajectory device status words (0 = clean)
    stats: dict | None = None                 # engine counters (matvecs, Krylov calls, SVD sweeps, kernel launches)
    segment_results: list = field(default_factory=list)   # SimulationProgram runs: one Result per segment (result.py:170-189)
    segment_index: int | None = None
    segment_type: str | None = None
    time_offset: float | None = None
