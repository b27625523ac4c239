"""The algebra behind the split final solve (zebrapose_b200/csrc/zp_finsplit.cu): EPnP's 52 sums over the inliers
(sum a_j a_k {1, x, y, x^2 + y^2}, sum a_j (X - c0)) are contractions of 40 raw moments T_f = sum f [X 1][X 1]^T, because the
barycentric coordinates are affine in the point, alpha = A [X 1]^T.  Checked here in numpy against the direct sums (the form
zp_final_cl_kernel accumulates), incl. a model whose origin lies far from the object and the pivot that handles it.
The GPU test test_final_solve_forms_identical compares the two kernels' poses."""
import numpy as np
import pytest


def _case(offset, pivot, seed=0, n=6000):
    rng = np.random.default_rng(seed)
    Xw = rng.uniform(-60, 60, (n, 3)) + offset
    R, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    Xc = Xw @ R.T + np.array([20., -30., 800.])
    fu, fv, uc, vc = 1066., 1067., 312., 241.
    u = fu * Xc[:, 0] / Xc[:, 2] + uc + rng.normal(0, .5, n)
    v = fv * Xc[:, 1] / Xc[:, 2] + vc + rng.normal(0, .5, n)
    Xw = Xw.astype(np.float32).astype(np.float64)
    u = u.astype(np.float32).astype(np.float64)
    v = v.astype(np.float32).astype(np.float64)
    # direct: centroid, scatter, control basis, alphas, sums (extended precision as the yardstick)
    c0 = Xw.mean(0)
    D = Xw - c0
    dc, uct = np.linalg.eigh(D.T @ D)
    cci = uct.T / np.sqrt(dc / n)[:, None]
    a123 = D @ cci.T
    al = np.concatenate([1 - a123.sum(1, keepdims=True), a123], 1).astype(np.longdouble)
    x, y = uc - u, vc - v
    fs = [np.ones(n), x, y, x * x + y * y]
    ref = [np.einsum('ij,ik,i->jk', al, al, f.astype(np.longdouble)).astype(np.float64) for f in fs]
    ref_w = (al.T @ D.astype(np.longdouble)).astype(np.float64)
    # moments relative to the pivot g (the kernel uses the crop's first 3D point)
    g = Xw[0] if pivot else np.zeros(3)
    P = np.concatenate([Xw - g, np.ones((n, 1))], 1)
    T = [np.einsum('ia,ib,i->ab', P, P, f) for f in fs]
    c0m = T[0][:3, 3] / T[0][3, 3]
    C = T[0][:3, :3] - np.outer(c0m, T[0][:3, 3])
    assert np.abs(C - D.T @ D).max() <= 1e-9 * np.abs(D.T @ D).max()
    A = np.zeros((4, 4))
    A[1:, :3] = cci
    A[1:, 3] = -cci @ c0m
    A[0] = np.array([0, 0, 0, 1.]) - A[1:].sum(0)
    got = [A @ Tf @ A.T for Tf in T]
    got_w = A @ (T[0][:, :3] - np.outer(T[0][:, 3], c0m))
    err = max(np.abs(a - b).max() / np.abs(b).max() for a, b in zip(got, ref))
    err_w = np.abs(got_w - ref_w).max() / np.abs(ref_w).max()
    return err, err_w


@pytest.mark.parametrize("offset", [(0, 0, 0), (30, -20, 40)])
def test_contractions_of_raw_moments_equal_epnp_sums(offset):
    err, err_w = _case(np.array(offset, float), pivot=True)
    assert err < 1e-12 and err_w < 1e-12, (err, err_w)


def test_pivot_keeps_the_cancellation_at_object_scale():
    """model origin 10 object sizes away: without the pivot the raw moments lose 3-4 digits, with it they do not"""
    far = np.array([500., -300., 1000.])
    e_no, w_no = _case(far, pivot=False)
    e_yes, w_yes = _case(far, pivot=True)
    assert e_yes < 1e-12 and w_yes < 1e-12, (e_yes, w_yes)
    assert e_no > 10 * e_yes
