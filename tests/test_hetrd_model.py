"""-m "not gpu": numpy restatement of the algebra behind csrc/yb_hetrd.cuh (blocked Hermitian tridiagonalisation with zlatrd panels,
reflectors stored in the rows of G, eigenvector rows E^H = Z^T Qh^H applied from the right).  It pins the conventions the CUDA kernels
implement -- conjugations, which triangle is read, where v_j lives, the order of the reflectors -- against numpy.linalg on the CPU; the
-m gpu tests (test_heev_rows_*) check the kernels themselves."""
import numpy as np
import pytest


def larfg(x):
    """zlarfg: v (v[0] = 1), tau, beta with (I - tau v v^H)^H x = beta e_1."""
    alpha = x[0]
    xn2 = np.vdot(x[1:], x[1:]).real
    if xn2 == 0 and alpha.imag == 0:
        v = np.zeros_like(x)
        v[0] = 1
        return v, 0j, alpha.real
    beta = -np.copysign(np.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xn2), alpha.real)
    tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
    v = x / (alpha - beta)
    v[0] = 1
    return v, tau, beta


def hetrd_blocked(G, NB):
    """k_hetrd_blocked: per panel the trailing block is only read (A0); V, W hold the rank-2 corrections; one update per panel."""
    G = G.copy()
    n = G.shape[0]
    d, e, tau = np.zeros(n), np.zeros(n), np.zeros(n, complex)
    for j0 in range(0, n - 1, NB):
        nb = min(NB, n - 1 - j0)
        V, W = np.zeros((n, nb), complex), np.zeros((n, nb), complex)
        for i in range(nb):
            j = j0 + i
            a = np.conj(G[j, j:]).copy()                               # column j through row j (Hermitian storage)
            a -= V[j:, :i] @ np.conj(W[j, :i]) + W[j:, :i] @ np.conj(V[j, :i])
            d[j] = a[0].real
            v, t, beta = larfg(a[1:])
            e[j], tau[j] = beta, t
            S = slice(j + 1, n)
            p = np.conj(G[S, S]).T @ v                                 # column sums of conj(G): (A0 v)[c] = sum_r conj(G[r][c]) v[r]
            p -= V[S, :i] @ (np.conj(W[S, :i]).T @ v) + W[S, :i] @ (np.conj(V[S, :i]).T @ v)
            w = t * p
            w = w + (-0.5 * t * np.vdot(w, v)) * v
            V[S, i], W[S, i] = v, w
            G[j, : j + 1] = 0
            G[j, j + 1:] = v                                           # row j <- v_j, zero-padded on the left
        T0 = j0 + nb
        G[T0:, T0:] -= V[T0:] @ np.conj(W[T0:]).T + W[T0:] @ np.conj(V[T0:]).T
    d[n - 1] = G[n - 1, n - 1].real
    G[n - 1, :] = 0
    return G, d, e, tau


def ormtr_rows(M, G, tau):
    """k_ormtr_rows: M <- M H_{n-2}^H ... H_0^H with H_j = I - tau_j v_j v_j^H, v_j = row j of G."""
    M = M.astype(complex).copy()
    for j in range(G.shape[0] - 2, -1, -1):
        v = G[j]
        M -= np.conj(tau[j]) * np.outer(M @ v, np.conj(v))
    return M


@pytest.mark.parametrize("n,NB", [(2, 16), (7, 3), (33, 8), (64, 16), (50, 16), (17, 16)])
def test_blocked_tridiagonalisation_and_row_back_transformation(n, NB):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n + 3)) + 1j * rng.standard_normal((n, n + 3))
    A *= np.logspace(0, -5, n + 3)[None, :]
    G0 = A @ A.conj().T
    G, d, e, tau = hetrd_blocked(G0, NB)
    T = np.diag(d) + np.diag(e[: n - 1], 1) + np.diag(e[: n - 1], -1)
    lam, Z = np.linalg.eigh(T)
    E = ormtr_rows(Z.T, G, tau).conj().T                               # rows of the result are E^H
    scale = np.abs(lam).max()
    assert np.abs(G0 @ E - E * lam[None, :]).max() < 1e-13 * scale
    assert np.abs(E.conj().T @ E - np.eye(n)).max() < 1e-13
    assert np.abs(np.sort(lam) - np.linalg.eigvalsh(G0)).max() < 1e-13 * scale
