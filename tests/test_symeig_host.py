"""CPU: the 12x12 symmetric eigen-solver of the device EPnP core (Householder + implicit QL, zp_epnp.cuh) compiled for
the host and checked against numpy.linalg.eigh: eigenvalues, residual, orthogonality, rank-deficient Gram matrices
(the 5-point EPnP case: M^T M of a 10x12 M has a 2-dimensional null space)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "symeig_host")
    src = os.path.join(ROOT, "tests", "native", "symeig_host.cu")
    r = subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", out, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def _run(exe, mats, mode="full"):
    txt = "\n".join(" ".join("%.17g" % v for v in A.ravel()) for A in mats) + "\n"
    out = subprocess.run([exe, mode], input=txt, capture_output=True, text=True, check=True).stdout
    a = np.array([[float(v) for v in line.split()] for line in out.strip().split("\n")])
    if mode == "small4":
        return a[:, :4], a[:, 4:].reshape(-1, 4, 12)
    return a[:, :12], a[:, 12:].reshape(-1, 12, 12)


def _cases():
    rng = np.random.default_rng(5)
    mats = []
    for _ in range(20):                                   # dense random symmetric
        A = rng.normal(size=(12, 12)); mats.append(A + A.T)
    for rows in (8, 10, 10, 10, 11, 12, 40):              # Gram matrices (rank-deficient below 12 rows), pixel-like scales
        for _ in range(6):
            M = rng.normal(size=(rows, 12)) * np.array([1000.0, 1000.0, 300.0] * 4)
            mats.append(M.T @ M)
    mats.append(np.diag(np.arange(12.0)))                 # already diagonal
    T = np.diag(np.arange(1.0, 13.0)) + np.diag(np.ones(11), 1) + np.diag(np.ones(11), -1)
    mats.append(T)                                        # already tridiagonal
    mats.append(np.zeros((12, 12)))
    mats.append(np.ones((12, 12)))                        # rank one, 11-fold zero eigenvalue
    return mats


def test_symeig12_matches_eigh(exe):
    mats = _cases()
    d, Z = _run(exe, mats)
    for A, w, V in zip(mats, d, Z):
        nrm = max(np.abs(A).sum(1).max(), 1e-300)
        ref = np.linalg.eigvalsh(A)
        assert np.allclose(np.sort(w), ref, rtol=0, atol=5e-14 * nrm)
        assert np.abs(V.T @ V - np.eye(12)).max() < 5e-14
        assert np.abs(A @ V - V * w).max() < 1e-13 * nrm


def test_symeig12_nan_terminates(exe):
    A = np.full((12, 12), np.nan)
    d, Z = _run(exe, [A])                                 # must return (garbage allowed), not hang
    assert d.shape == (1, 12)


def test_smallest4_matches_eigh(exe):
    """zp_smallest4_12 (bisection + inverse iteration, what the kernels run): eigenvalues, residuals, orthonormality, and
    -- where eigenvalues coincide (rank-deficient Gram matrices) -- the right invariant subspace"""
    mats = [A for A in _cases() if np.abs(A).max() > 0]
    lam, V = _run(exe, mats, "small4")
    for A, w, v in zip(mats, lam, V):
        nrm = np.abs(A).sum(1).max()
        ew, ev = np.linalg.eigh(A)
        assert np.allclose(w, ew[:4], rtol=0, atol=2e-13 * nrm)
        assert np.abs(v @ v.T - np.eye(4)).max() < 1e-10
        assert np.abs(A @ v.T - v.T * w).max() < 2e-12 * nrm
        # eigenvalues 0..3 separated from the rest: the span must be the span of eigh's four vectors
        if ew[4] - ew[3] > 1e-6 * nrm:
            P = ev[:, :4] @ ev[:, :4].T
            assert np.abs(v @ P - v).max() < 1e-7
