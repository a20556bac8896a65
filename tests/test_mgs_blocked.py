"""The blocked modified Gram-Schmidt of the device Krylov loop (csrc/solver.cu k_mgs_pass / mgs_betas), restated in
numpy: four basis vectors per pass, coefficients from the dot products of the pass and the Gram entries of its block,
    beta_k = (w,u_k) - sum_{l<k} beta_l (u_l,u_k),
against the reference's sweep (solgmr.f:224-244: one dot product and one subtraction per basis vector).  Linearity of
the dot product makes the two identical in exact arithmetic whether or not the basis is orthogonal; in floating point
they agree to the rounding of the sums -- the property the GPU path's HBrg <= 1e-8 parity rests on."""
import numpy as np
import pytest


def mgs_reference(w, U):
    """solgmr.f:224-244 + :246-256: returns the Hessenberg column (betas, norm) and the new unit vector"""
    w = w.copy()
    h = []
    for u in U:
        b = float(w @ u)
        h.append(b)
        w -= b * u
    nrm = float(np.sqrt(w @ w))
    return np.array(h + [nrm]), w / nrm


def mgs_blocked(w, U, B=4):
    """k_mgs_pass: pass p subtracts block p-1 and reduces [dots(4) | gram(6) | (w,w)] of block p"""
    w = w.copy()
    n = len(U)
    nblk = (n + B - 1) // B
    h = []
    beta, prev = None, None
    red = None
    for p in range(nblk + 1):
        if p > 0:
            blk = U[B * (p - 1):B * p]
            d, G = red
            beta = np.zeros(len(blk))
            for k in range(len(blk)):                     # mgs_betas
                beta[k] = d[k] - sum(beta[l] * G[l][k] for l in range(k))
            for k, u in enumerate(blk):                   # same order of subtractions as the sweep
                w = w - beta[k] * u
            h += list(beta)
        if p < nblk:
            blk = U[B * p:B * (p + 1)]
            d = [float(w @ u) for u in blk]
            G = [[float(a @ b) for b in blk] for a in blk]
            red = (d, G)
        else:
            nrm = float(np.sqrt(w @ w))
    return np.array(h + [nrm]), w / nrm


@pytest.mark.parametrize("n", [1, 3, 4, 5, 8, 13, 30])
def test_blocked_equals_sweep_on_an_orthonormal_basis(n):
    rng = np.random.default_rng(n)
    Q, _ = np.linalg.qr(rng.standard_normal((2000, n)))
    U = [Q[:, k].copy() for k in range(n)]
    w = rng.standard_normal(2000)
    h0, u0 = mgs_reference(w, U)
    h1, u1 = mgs_blocked(w, U)
    assert np.max(np.abs(h0 - h1)) < 1e-13 * np.linalg.norm(h0)
    assert np.linalg.norm(u0 - u1) < 1e-12


def test_blocked_equals_sweep_when_the_basis_has_lost_orthogonality():
    """the Gram terms are what makes it MODIFIED Gram-Schmidt: with a visibly non-orthogonal basis the classical
    variant (beta_k = (w,u_k)) is off by the size of the overlaps, the blocked recurrence is not"""
    rng = np.random.default_rng(7)
    Q, _ = np.linalg.qr(rng.standard_normal((500, 8)))
    U = [Q[:, k] + 1e-3 * rng.standard_normal(500) for k in range(8)]
    U = [u / np.linalg.norm(u) for u in U]
    w = rng.standard_normal(500)
    h0, u0 = mgs_reference(w, U)
    h1, u1 = mgs_blocked(w, U)
    assert np.max(np.abs(h0 - h1)) < 1e-12 * np.linalg.norm(h0)
    classical = np.array([float(w @ u) for u in U])
    assert np.max(np.abs(classical - h0[:-1])) > 1e-5          # the difference the Gram terms account for


def test_a_whole_arnoldi_loop_gives_the_same_hessenberg():
    """(ten steps: the residual is still ~1e-5 of its start; beyond convergence the late Krylov vectors are rounding
    noise in EITHER variant and their Hessenberg entries carry no information)"""
    rng = np.random.default_rng(3)
    n, m = 300, 10
    A = np.eye(n) + 0.3 * rng.standard_normal((n, n)) / np.sqrt(n)
    r = rng.standard_normal(n)
    out = []
    for mgs in (mgs_reference, mgs_blocked):
        U = [r / np.linalg.norm(r)]
        H = np.zeros((m + 1, m))
        for k in range(m):
            h, u = mgs(A @ U[k], U)
            H[:k + 2, k] = h
            U.append(u)
        out.append(H)
    assert np.max(np.abs(out[0] - out[1])) < 1e-11
