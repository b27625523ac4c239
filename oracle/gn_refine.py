"""Gauss-Newton polish of the pixel reprojection error over the inliers (test infrastructure): the float64 twin of
zp_final_kernel's `final="epnp+gn"` stage (north_star's "batched Gauss-Newton refine on the inliers").  The reference has
no such step -- cv2.solvePnPRansac ends with EPnP on the inliers (CNN_output_to_pose.py:155-157) -- so this twin pins the
device arithmetic, and cv2.solvePnPRefineLM (converged Levenberg-Marquardt on the same points) is the independent
cross-check of what it converges to (tests/test_gpu_ransac.py::test_gn_refine_against_twin)."""
import numpy as np


def _rodrigues(w):
    th = float(np.linalg.norm(w))
    if th <= 1e-300:
        return np.eye(3)
    k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * Kx


def gn_refine(R, t, pw, uv, K, iters=5):
    """R 3x3, t 3, pw [n,3], uv [n,2], K 3x3 -> (R, t) after `iters` Gauss-Newton steps with the left-multiplicative update
    cam = exp(w) (R X) + t + dt, normal equations solved by Cholesky (a step whose matrix is not positive definite is skipped)."""
    R = np.array(R, np.float64)
    t = np.array(t, np.float64).reshape(3)
    pw = np.asarray(pw, np.float64)
    uv = np.asarray(uv, np.float64)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    for _ in range(iters):
        P = pw @ R.T
        xc, yc, zc = P[:, 0] + t[0], P[:, 1] + t[1], P[:, 2] + t[2]
        iz = 1.0 / zc
        ru = fx * xc * iz + cx - uv[:, 0]
        rv = fy * yc * iz + cy - uv[:, 1]
        ju = np.stack([fx * iz, np.zeros_like(iz), -fx * xc * iz * iz], 1)
        jv = np.stack([np.zeros_like(iz), fy * iz, -fy * yc * iz * iz], 1)
        px, py, pz = P[:, 0], P[:, 1], P[:, 2]

        def full(j):
            return np.stack([j[:, 1] * (-pz) + j[:, 2] * py, j[:, 0] * pz + j[:, 2] * (-px), j[:, 0] * (-py) + j[:, 1] * px,
                             j[:, 0], j[:, 1], j[:, 2]], 1)
        Ju, Jv = full(ju), full(jv)
        A = Ju.T @ Ju + Jv.T @ Jv
        g = -(Ju.T @ ru + Jv.T @ rv)
        try:
            L = np.linalg.cholesky(A)
        except np.linalg.LinAlgError:
            continue
        d = np.linalg.solve(L.T, np.linalg.solve(L, g))
        R = _rodrigues(d[:3]) @ R
        t = t + d[3:]
    return R, t
