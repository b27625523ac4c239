"""Control-flow emulation of cv2.solvePnPRansac(EPnP) (test infrastructure).

cv::RNG(0xFFFFFFFFFFFFFFFF) multiply-with-carry sampler, 5-point minimal sets, float32 squared reprojection
error <= thr^2, strictly-greater model update, RANSACUpdateNumIters(0.99), final EPnP on the inliers.
The survey verified identical inlier arrays versus cv2.solvePnPRansac on 5 problems; tests/test_oracle_ransac.py
repeats that check.  `solver` is pluggable: cv2.solvePnP (black-box reference arithmetic) or oracle.epnp.
"""
import math
import numpy as np
import cv2

DBL_MIN = 2.2250738585072014e-308


class CvRNG:
    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state or 0xFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def update_iters(p, ep, m, maxit):
    """RANSACUpdateNumIters"""
    p = max(p, 0.0); p = min(p, 1.0)
    ep = max(ep, 0.0); ep = min(ep, 1.0)
    num = max(1.0 - p, DBL_MIN)
    den = 1.0 - (1.0 - ep) ** m
    if den < DBL_MIN:
        return 0
    num, den = math.log(num), math.log(den)
    return maxit if (den >= 0 or -num >= maxit * (-den)) else int(round(num / den))


def sample_lists(n, iters=150, m=5):
    """The `iters` minimal-sample index lists cv2 would draw for n correspondences (depends on n only)."""
    rng = CvRNG()
    out = np.empty((iters, m), np.int32)
    for it in range(iters):
        idx = []
        for _ in range(m):
            j = rng.uniform(0, n)
            while j in idx:
                j = rng.uniform(0, n)
            idx.append(j)
        out[it] = idx
    return out


def cv2_solver(obj, img, K):
    ok, rv, tv = cv2.solvePnP(obj, img, K, None, flags=cv2.SOLVEPNP_EPNP)
    if not ok:
        return None
    return cv2.Rodrigues(rv)[0], tv.reshape(3)


def score_pose(obj_f32, img_f32, K, R, t, thr=2.0):
    """PnPRansacCallback::computeError: projectPoints in float64 -> float32, float32 squared distance."""
    rv = cv2.Rodrigues(np.asarray(R, np.float64))[0]
    proj = cv2.projectPoints(obj_f32, rv, np.asarray(t, np.float64).reshape(3, 1), K, None)[0].reshape(-1, 2)
    proj = proj.astype(np.float32)
    d = proj - img_f32
    err = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)
    return err <= np.float32(thr * thr), err


def replay_select(counts, n, m=5, iters=150, conf=0.99):
    """Sequential cv2 model-update rule replayed on pre-computed inlier counts.  Returns (best index or -1,
    number of iterations cv2 would have run)."""
    niters = max(iters, 1)
    maxgood = 0
    best = -1
    it = 0
    while it < niters and it < len(counts):
        good = int(counts[it])
        if good > max(maxgood, m - 1):
            best, maxgood = it, good
            niters = update_iters(conf, (n - good) / n, m, niters)
        it += 1
    return best, it


def solve_pnp_ransac(obj_f32, img_f32, K, thr=2.0, iters=150, conf=0.99, m=5, solver=cv2_solver,
                     final_solver=None, log=None):
    """Emulated cv2.solvePnPRansac.  Returns (ok, R, t, inlier idx, info dict)."""
    n = len(obj_f32)
    samples = sample_lists(n, iters, m)
    niters = max(iters, 1)
    maxgood = 0
    best = None
    it = 0
    counts = []
    while it < niters:
        idx = samples[it]
        it += 1
        sol = solver(obj_f32[idx], img_f32[idx], K)
        if sol is None:
            counts.append(0)
            continue
        mask, _ = score_pose(obj_f32, img_f32, K, sol[0], sol[1], thr)
        good = int(mask.sum())
        counts.append(good)
        if log is not None:
            log.append((idx.copy(), sol[0].copy(), sol[1].copy(), good))
        if good > max(maxgood, m - 1):
            best, maxgood = (mask, it - 1), good
            niters = update_iters(conf, (n - good) / n, m, niters)
    info = dict(iters_run=it, counts=counts, samples=samples)
    if best is None:
        return False, np.eye(3), np.zeros(3), np.zeros(0, np.int64), info
    inl = np.nonzero(best[0])[0]
    info["best"] = best[1]
    fs = final_solver or solver
    sol = fs(obj_f32[inl].astype(np.float64), img_f32[inl].astype(np.float64), K)
    if sol is None:
        return False, np.eye(3), np.zeros(3), inl, info
    return True, sol[0], sol[1], inl, info
