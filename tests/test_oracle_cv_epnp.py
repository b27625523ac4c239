"""CPU: oracle/cv_epnp.c (operation-by-operation restatement of OpenCV's EPnP arithmetic) pinned against cv2 itself, and
the product's exact solver (zebrapose_b200/csrc/zp_cvepnp.cuh, host build with emulated lanes) pinned against both.

OpenCV is the un-vendored library behind the reference's cv2.solvePnPRansac call (CNN_output_to_pose.py:155-157); the pin
is opencv-python-headless 4.13 as installed in this image."""
import ctypes as C
import os
import subprocess

import cv2
import numpy as np
import pytest

from oracle import cv_epnp, cvransac, decode
from workloads import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank_deficient(rng, rows, n=12):
    M = rng.normal(size=(rows, n)) * np.array([500, 500, 100] * (n // 3))
    return M


def test_stages_bit_identical_to_cv2():
    rng = np.random.default_rng(0)
    for _ in range(20):
        M = _rank_deficient(rng, 10)
        mtm = cv2.mulTransposed(M, True)
        assert np.array_equal(mtm, cv_epnp.mul_transposed(M))                    # sequential sums, no FMA
        w, u, vt = cv2.SVDecomp(mtm)                                             # rank 10: two null vectors set by rounding
        W, Ut, Vt = cv_epnp.svd_square(mtm)
        assert np.array_equal(w.ravel(), W) and np.array_equal(u.T, Ut) and np.array_equal(vt, Vt)
    for _ in range(50):
        A = rng.normal(size=(3, 3)) * 30
        w, u, vt = cv2.SVDecomp(A)
        W, Ut, Vt = cv_epnp.svd_square(A)
        assert np.array_equal(w.ravel(), W) and np.array_equal(u.T, Ut) and np.array_equal(vt, Vt)
        ok, inv = cv2.invert(A, flags=cv2.DECOMP_SVD)
        assert np.array_equal(inv, cv_epnp.invert3_svd(A))
    for nc in (3, 4, 5):
        for _ in range(30):
            A, b = rng.normal(size=(6, nc)), rng.normal(size=(6, 1))
            ok, x = cv2.solve(A, b, flags=cv2.DECOMP_SVD)
            assert np.array_equal(x.ravel(), cv_epnp.solve_svd(A, b))


def _crop_lists(seed, bitflip=0.0):
    tab, nrm, _ = synth.make_dict(16, 0, 50.0, 0.0)
    c = synth.make_crop(tab, nrm, seed, bitflip=bitflip)
    uv, xyz, _ = decode.decode_crop(c["mask"], c["bits"].astype(np.float64), c["bbox"], 128, tab)
    return c, uv, xyz


@pytest.mark.parametrize("m", [4, 5, 6, 8])
def test_minimal_sample_poses_equal_cv2(m):
    """whole poses of outlier-bearing minimal samples: rvec and tvec bit-identical to cv2.solvePnP's"""
    n_checked = 0
    for seed, flip in ((500, 0.0), (501, 0.02)):
        c, uv, xyz = _crop_lists(seed, flip)
        for idx in cvransac.sample_lists(len(uv), 60, m):
            ok, rv, tv = cv2.solvePnP(xyz[idx], uv[idx], c["K"], None, flags=cv2.SOLVEPNP_EPNP)
            R, t = cv_epnp.epnp(xyz[idx], uv[idx], c["K"])
            if not np.all(np.isfinite(rv)):
                assert not np.all(np.isfinite(R))
                continue
            assert np.array_equal(cv2.Rodrigues(R)[0], rv) and np.array_equal(t, tv.ravel())
            n_checked += 1
    assert n_checked >= 100


def test_large_n_matches_cv2():
    for n in (50, 3000):
        c, uv, xyz = _crop_lists(502)
        ok, rv, tv = cv2.solvePnP(xyz[:n], uv[:n], c["K"], None, flags=cv2.SOLVEPNP_EPNP)
        R, t = cv_epnp.epnp(xyz[:n], uv[:n], c["K"])
        assert np.array_equal(cv2.Rodrigues(R)[0], rv) and np.array_equal(t, tv.ravel())


def test_ransac_with_oracle_solver_equals_cv2():
    """the control-flow emulation with the restated solver reproduces cv2.solvePnPRansac: same inliers, same pose"""
    for seed in (510, 511):
        c, uv, xyz = _crop_lists(seed, 0.02 if seed == 511 else 0.0)
        ok, rv, tv, inl = cv2.solvePnPRansac(xyz, uv, c["K"], None, reprojectionError=2, iterationsCount=150,
                                             flags=cv2.SOLVEPNP_EPNP)
        ok2, R2, t2, inl2, info = cvransac.solve_pnp_ransac(xyz, uv, c["K"], solver=cv_epnp.solver)
        assert np.array_equal(inl.ravel(), inl2)
        assert np.array_equal(cv2.Rodrigues(R2)[0], rv) and np.array_equal(t2, tv.ravel())


# ---------------------------------------------------------------------------------------------------------------------
# the product source, compiled for the host
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("native") / "cvepnp_host.so")
    src = os.path.join(ROOT, "tests", "native", "cvepnp_host.cpp")
    r = subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(so)


def _host_epnp(host, corr, idx, K):
    K4 = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]])
    pose, st = np.zeros(12), (C.c_int * 1)()
    idx = np.ascontiguousarray(idx, np.int32)
    vp = C.c_void_p
    host.cve_host_epnp(corr.ctypes.data_as(vp), corr.shape[1], idx.ctypes.data_as(vp), len(idx), K4.ctypes.data_as(vp),
                       pose.ctypes.data_as(vp), None, st)
    return pose, list(st)


@pytest.mark.parametrize("m", [4, 5, 6, 7, 8])
def test_product_solver_bit_identical_to_oracle(host, m):
    """stage A / C serial, stage B on six emulated lanes in the wave-front pair order: identical bits to the serial
    restatement (and so to cv2)"""
    tot, steps = 0, []
    for seed, flip in ((700, 0.0), (701, 0.02)):
        c, uv, xyz = _crop_lists(seed, flip)
        corr = np.ascontiguousarray(np.concatenate([uv.T, xyz.T]).astype(np.float32))
        for idx in cvransac.sample_lists(len(uv), 100, m):
            R, t = cv_epnp.epnp(xyz[idx], uv[idx], c["K"])
            pose, st = _host_epnp(host, corr, idx, c["K"])
            want = np.concatenate([R.ravel(), t])
            assert np.array_equal(want, pose, equal_nan=True), (m, seed, idx)
            tot += 1
            steps.append(sum(st))
    assert tot == 200
    assert np.mean(steps) < 110          # 12x12: the serial pair order needs ~400 pair steps


def test_product_solver_degenerate_inputs(host):
    """all object points identical / collinear / zero: exactly-zero singular values take OpenCV's pseudo-random-vector
    branch; whatever comes out (NaN included) must be what the restatement produces"""
    rng = np.random.default_rng(3)
    K = synth.LM_K
    cases = []
    uv = np.trunc(rng.uniform(100, 400, size=(6, 2)))
    cases.append((np.zeros((6, 3)), uv))
    cases.append((np.tile(rng.normal(size=(1, 3)) * 30, (6, 1)), uv))
    line = np.outer(np.arange(6.0), [1.0, 2.0, -1.0]) + 5
    cases.append((line, uv))
    plane = rng.normal(size=(6, 3)) * 30
    plane[:, 2] = 0
    cases.append((plane, uv))
    for pw, uv in cases:
        pw32 = pw.astype(np.float32)
        corr = np.ascontiguousarray(np.concatenate([uv.T, pw32.T]).astype(np.float32))
        uv32 = uv.astype(np.float32)
        R, t = cv_epnp.epnp(pw32, uv32, K)
        pose, _ = _host_epnp(host, corr, np.arange(6), K)
        assert np.array_equal(np.concatenate([R.ravel(), t]), pose, equal_nan=True)
        ok, rv, tv = cv2.solvePnP(pw32, uv32, K, None, flags=cv2.SOLVEPNP_EPNP)
        assert np.array_equal(tv.ravel(), t, equal_nan=True)
