"""ctypes front of oracle/cv_epnp.c (test infrastructure): the operation-by-operation restatement of the arithmetic of
cv2.solvePnP(SOLVEPNP_EPNP), the solver behind the reference's cv2.solvePnPRansac call
(zebrapose/binary_code_helper/CNN_output_to_pose.py:155-157).  Pinned bit for bit against cv2 4.13 by
tests/test_oracle_cv_epnp.py.  `make -C oracle` (also run by __graft_entry__.build()) builds the library."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libcv_epnp.so")
_lib = None
_P = C.c_void_p


def build():
    src = os.path.join(HERE, "cv_epnp.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(_P)


def epnp(pw, uv, K, stage=None, debug=False):
    """pw [n,3], uv [n,2] pixels, K 3x3 -> (R 3x3, t 3) float64 exactly as cv2.solvePnP(pw, uv, K, None,
    flags=SOLVEPNP_EPNP) computes them (R before its Rodrigues round trip).  stage: how cv2 stages the image points --
    1 float32 normalised coordinates (float32 image points: the RANSAC hypotheses), 2 float64 (float64 image points: the
    final solve on the inliers), 0 none; None = by the dtype of uv, as cv2 does."""
    if stage is None:
        stage = 1 if np.asarray(uv).dtype == np.float32 else 2
    pw = np.ascontiguousarray(pw, np.float64)
    uv = np.ascontiguousarray(uv, np.float64)
    n = len(pw)
    K = np.asarray(K, np.float64)
    K4 = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]])
    scr, R, t, dbg = np.zeros(9 * n), np.zeros(9), np.zeros(3), np.zeros(171)
    N = lib().zpo_cv_epnp(_p(pw), _p(uv), n, _p(K4), int(stage), _p(scr), _p(R), _p(t), _p(dbg))
    if debug:
        return R.reshape(3, 3), t, dict(N=N, ut=dbg[:144].reshape(12, 12), d=dbg[144:156],
                                        betas=dbg[156:168].reshape(3, 4), rep=dbg[168:171])
    return R.reshape(3, 3), t


def solver(obj, img, K):
    """drop-in for oracle.cvransac's `solver` argument"""
    R, t = epnp(obj, img, K)
    if not (np.all(np.isfinite(R)) and np.all(np.isfinite(t))):
        return None
    return R, t


def svd_square(A):
    """(w, Ut, Vt) of a square matrix as the small-matrix Jacobi SVD returns them"""
    A = np.ascontiguousarray(A, np.float64)
    n = A.shape[0]
    w, ut, vt = np.zeros(n), np.zeros((n, n)), np.zeros((n, n))
    lib().zpo_svd_square(_p(A), n, _p(w), _p(ut), _p(vt))
    return w, ut, vt


def mul_transposed(M):
    M = np.ascontiguousarray(M, np.float64)
    out = np.zeros((M.shape[1], M.shape[1]))
    lib().zpo_mul_transposed(_p(M), M.shape[0], M.shape[1], _p(out))
    return out


def invert3_svd(A):
    A = np.ascontiguousarray(A, np.float64)
    out = np.zeros((3, 3))
    lib().zpo_invert3_svd(_p(A), _p(out))
    return out


def solve_svd(A, b):
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64).ravel()
    x = np.zeros(A.shape[1])
    lib().zpo_solve_svd(_p(A), A.shape[0], A.shape[1], _p(b), _p(x))
    return x


def census(samples):
    """average arithmetic operations of the Jacobi SVDs per solve over `samples` = [(pw, uv, K), ...]: dict(ops_n12, ops_other,
    rotated_n12, skipped_n12).  (OpenCV also rotates a 12 x 12 Vt for the n = 12 problem, which the device kernel does not
    need and the census therefore leaves out there: 6 n per rotated pair.)"""
    L = lib()
    L.zpo_census_reset()
    for pw, uv, K in samples:
        epnp(pw, uv, K)
    out = np.zeros(4)
    L.zpo_census_read(_p(out))
    n = max(1, len(samples))
    return dict(ops_n12=(out[0] - 72.0 * out[2]) / n, ops_other=out[1] / n, rotated_n12=out[2] / n, skipped_n12=out[3] / n)
