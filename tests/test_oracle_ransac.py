"""CPU: the RANSAC / EPnP oracle against cv2 itself (the un-vendored arithmetic behind the reference path)."""
import cv2
import numpy as np
import pytest

from oracle import cvransac, decode, epnp, metrics
from workloads import synth


def _problem(seed, n, noise=True, sigma=0.0):
    rng = np.random.default_rng(seed)
    pw = (rng.normal(size=(n, 3)) * 40).astype(np.float32)
    R, t = synth.random_pose(rng)
    P = (R @ pw.T).T + t
    uv = (synth.LM_K @ P.T).T
    uv = uv[:, :2] / uv[:, 2:]
    if noise:
        uv = np.trunc(uv + rng.normal(size=uv.shape) * sigma)
    return pw, uv.astype(np.float32), R, t


@pytest.mark.parametrize("sigma", [0.0, 30.0])
@pytest.mark.parametrize("n", [6, 8, 50, 2000])
def test_epnp_matches_cv2(n, sigma):
    """incl. heavy pixel noise (outlier-bearing samples): pins the pixel-unit formulation cv2 uses"""
    for s in range(8):
        pw, uv, _, _ = _problem(100 * n + s, n, sigma=sigma)
        Rc, tc = cvransac.cv2_solver(pw, uv, synth.LM_K)
        Ro, to = epnp.epnp(pw, uv, synth.LM_K, f32_inputs=True)
        assert metrics.rot_err_deg(Rc, Ro) < 1e-3, (n, s)
        assert metrics.trans_err(tc, to) < 1e-2, (n, s)


def test_sample_lists_match_cv2_inliers():
    """emulated control flow (RNG replay, strictly-greater update, adaptive stop, final EPnP) == cv2.solvePnPRansac"""
    tab, nrm, _ = synth.make_dict(16, 0, 50.0, 0.0)
    for s in range(3):
        c = synth.make_crop(tab, nrm, 500 + s)
        uv, xyz, _ = decode.decode_crop(c["mask"], c["bits"].astype(np.float64), c["bbox"], 128, tab)
        ok, rv, tv, inl = cv2.solvePnPRansac(xyz, uv, c["K"], None, reprojectionError=2, iterationsCount=150,
                                             flags=cv2.SOLVEPNP_EPNP)
        ok2, R2, t2, inl2, info = cvransac.solve_pnp_ransac(xyz, uv, c["K"])
        assert np.array_equal(inl.ravel(), inl2)
        assert metrics.rot_err_deg(cv2.Rodrigues(rv)[0], R2) < 1e-4
        best, it = cvransac.replay_select(info["counts"] + [0] * (150 - len(info["counts"])), len(uv))
        assert best == info["best"] and it == info["iters_run"]


def test_cvrng_first_values():
    r = cvransac.CvRNG()
    a = [r.next() for _ in range(3)]
    r2 = cvransac.CvRNG()
    assert a == [r2.next() for _ in range(3)] and all(0 <= v < 2 ** 32 for v in a)
    s = cvransac.sample_lists(100, 10, 5)
    assert s.shape == (10, 5) and all(len(set(row)) == 5 for row in s) and s.max() < 100
