"""float64 restatement of OpenCV's EPnP as cv2.solvePnP(..., flags=SOLVEPNP_EPNP) runs it (test infrastructure).

OpenCV is an un-vendored dependency of the reference (CNN_output_to_pose.py:155-157; opencv-python-headless
4.13.0.92 here).  Restated from the published algorithm (Lepetit/Moreno-Noguer/Fua 2009) plus the behaviours
of OpenCV's implementation that change the answer under noise, each verified against cv2 in this image:
  * M is built in PIXEL units with the real camera matrix (rows [a*fu, 0, a*(uc-u)] / [0, a*fv, a*(vc-v)]), so
    fu != fv weights the two image axes differently; a normalised-coordinate EPnP agrees on noise-free data but
    drifts by 1e-3..5e-2 deg on outlier-bearing samples because 5 Gauss-Newton steps are mid-transient there;
  * the PCA control points use the eigenvectors *with the signs* OpenCV's small-matrix one-sided Jacobi SVD
    returns (rows of the rotated A^T, normalised) -- a textbook eigh() differs by 0.1 deg under pixel noise;
  * three beta initialisations + 5 Gauss-Newton steps each, pose by Horn with "negate row 2" on det<0,
    best of three by mean reprojection distance.
Agreement with cv2.solvePnP(EPNP): <= 2e-5 deg / 4e-4 mm for n >= 6 (tests/test_oracle_epnp.py).  For n in
{4,5} M^T M has a 4-/2-dimensional null space whose basis (and hence cv2's own answer) is set by rounding
noise: parity unpinned there.
"""
import math
import numpy as np

_EPS = np.finfo(np.float64).eps * 10
_PAIRS = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]


def jacobi_svd_ut(A):
    """One-sided (Hestenes) Jacobi on the rows of A^T in OpenCV's pair order; returns (w desc, Ut) with
    Ut rows = left singular vectors = normalised rotated rows."""
    n = A.shape[0]
    At = np.array(A.T, dtype=np.float64, copy=True)
    W = (At * At).sum(1)
    for _ in range(max(n, 30)):
        changed = False
        for i in range(n - 1):
            for j in range(i + 1, n):
                a, b = W[i], W[j]
                p = float(At[i] @ At[j])
                if abs(p) <= _EPS * math.sqrt(a * b):
                    continue
                p *= 2
                beta = a - b
                gamma = math.hypot(p, beta)
                if beta < 0:
                    delta = (gamma - beta) * 0.5
                    s = math.sqrt(delta / gamma)
                    c = p / (gamma * s * 2)
                else:
                    c = math.sqrt((gamma + beta) / (gamma * 2))
                    s = p / (gamma * c * 2)
                t0 = c * At[i] + s * At[j]
                t1 = -s * At[i] + c * At[j]
                At[i], At[j] = t0, t1
                W[i], W[j] = float(t0 @ t0), float(t1 @ t1)
                changed = True
        if not changed:
            break
    W = np.sqrt((At * At).sum(1))
    for i in range(n - 1):              # selection sort, descending (stable w.r.t. OpenCV's swaps)
        j = i
        for k in range(i + 1, n):
            if W[j] < W[k]:
                j = k
        if i != j:
            W[[i, j]] = W[[j, i]]
            At[[i, j]] = At[[j, i]]
    for i in range(n):
        At[i] *= (1.0 / W[i]) if W[i] > 2.2250738585072014e-308 else 0.0
    return W, At


def _betas(L, rho):
    b = np.linalg.lstsq(L[:, [0, 1, 3, 6]], rho, rcond=None)[0]
    if b[0] < 0:
        r = math.sqrt(-b[0]); B1 = np.array([r, -b[1] / r, -b[2] / r, -b[3] / r])
    else:
        r = math.sqrt(b[0]); B1 = np.array([r, b[1] / r, b[2] / r, b[3] / r])
    b = np.linalg.lstsq(L[:, [0, 1, 2]], rho, rcond=None)[0]
    if b[0] < 0:
        B2 = np.array([math.sqrt(-b[0]), math.sqrt(-b[2]) if b[2] < 0 else 0.0, 0.0, 0.0])
    else:
        B2 = np.array([math.sqrt(b[0]), math.sqrt(b[2]) if b[2] > 0 else 0.0, 0.0, 0.0])
    if b[1] < 0:
        B2[0] = -B2[0]
    b = np.linalg.lstsq(L[:, [0, 1, 2, 3, 4]], rho, rcond=None)[0]
    if b[0] < 0:
        B3 = np.array([math.sqrt(-b[0]), math.sqrt(-b[2]) if b[2] < 0 else 0.0, 0.0, 0.0])
    else:
        B3 = np.array([math.sqrt(b[0]), math.sqrt(b[2]) if b[2] > 0 else 0.0, 0.0, 0.0])
    if b[1] < 0:
        B3[0] = -B3[0]
    B3[2] = b[3] / B3[0]
    return B1, B2, B3


def _gauss_newton(L, rho, b):
    b = b.copy()
    for _ in range(5):
        A = np.stack([2 * L[:, 0] * b[0] + L[:, 1] * b[1] + L[:, 3] * b[2] + L[:, 6] * b[3],
                      L[:, 1] * b[0] + 2 * L[:, 2] * b[1] + L[:, 4] * b[2] + L[:, 7] * b[3],
                      L[:, 3] * b[0] + L[:, 4] * b[1] + 2 * L[:, 5] * b[2] + L[:, 8] * b[3],
                      L[:, 6] * b[0] + L[:, 7] * b[1] + L[:, 8] * b[2] + 2 * L[:, 9] * b[3]], 1)
        r = rho - (L[:, 0] * b[0] * b[0] + L[:, 1] * b[0] * b[1] + L[:, 2] * b[1] * b[1] + L[:, 3] * b[0] * b[2]
                   + L[:, 4] * b[1] * b[2] + L[:, 5] * b[2] * b[2] + L[:, 6] * b[0] * b[3] + L[:, 7] * b[1] * b[3]
                   + L[:, 8] * b[2] * b[3] + L[:, 9] * b[3] * b[3])
        b = b + np.linalg.lstsq(A, r, rcond=None)[0]
    return b


def epnp(pw, uv, K, f32_inputs=True):
    """pw [n,3], uv [n,2] pixels, K 3x3 -> (R 3x3, t 3) float64.  `f32_inputs` is kept for call compatibility: the
    reference's image points are integers, exactly representable in float32, so cv2's float32 staging of them is a
    no-op here."""
    pw = np.asarray(pw, np.float64)
    uv = np.asarray(uv, np.float64)
    n = len(pw)
    fu, fv, uc, vc = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    xn = uv[:, 0]
    yn = uv[:, 1]
    c0 = pw.sum(0) / n
    d = pw - c0
    dc, uct = jacobi_svd_ut(d.T @ d)
    cws = np.zeros((4, 3))
    cws[0] = c0
    for i in range(3):
        cws[i + 1] = c0 + math.sqrt(dc[i] / n) * uct[i]
    CC = (cws[1:] - cws[0]).T
    CCi = np.linalg.pinv(CC)
    al = np.empty((n, 4))
    al[:, 1:] = d @ CCi.T
    al[:, 0] = 1.0 - al[:, 1] - al[:, 2] - al[:, 3]
    M = np.zeros((2 * n, 12))
    for j in range(4):
        M[0::2, 3 * j] = al[:, j] * fu
        M[0::2, 3 * j + 2] = al[:, j] * (uc - xn)
        M[1::2, 3 * j + 1] = al[:, j] * fv
        M[1::2, 3 * j + 2] = al[:, j] * (vc - yn)
    _, ut = jacobi_svd_ut(M.T @ M)
    v = [ut[11], ut[10], ut[9], ut[8]]
    dv = np.zeros((4, 6, 3))
    for k in range(4):
        for r, (a, b) in enumerate(_PAIRS):
            dv[k, r] = v[k][3 * a:3 * a + 3] - v[k][3 * b:3 * b + 3]
    L = np.zeros((6, 10))
    for r in range(6):
        q = dv[:, r]
        L[r] = [q[0] @ q[0], 2 * q[0] @ q[1], q[1] @ q[1], 2 * q[0] @ q[2], 2 * q[1] @ q[2], q[2] @ q[2],
                2 * q[0] @ q[3], 2 * q[1] @ q[3], 2 * q[2] @ q[3], q[3] @ q[3]]
    rho = np.array([((cws[a] - cws[b]) ** 2).sum() for a, b in _PAIRS])
    pw0 = pw.sum(0) / n
    best = None
    for B in _betas(L, rho):
        if not np.all(np.isfinite(B)):
            continue
        b = _gauss_newton(L, rho, B)
        ccs = sum(b[k] * v[k] for k in range(4)).reshape(4, 3)
        pcs = al @ ccs
        if pcs[0, 2] < 0:
            ccs, pcs = -ccs, -pcs
        pc0 = pcs.sum(0) / n
        ABt = (pcs - pc0).T @ (pw - pw0)
        U, _, Vt = np.linalg.svd(ABt)
        R = U @ Vt
        if np.linalg.det(R) < 0:
            R[2] = -R[2]
        t = pc0 - R @ pw0
        P = pw @ R.T + t
        iz = 1.0 / P[:, 2]
        err = np.sqrt((xn - (uc + fu * P[:, 0] * iz)) ** 2 + (yn - (vc + fv * P[:, 1] * iz)) ** 2).sum() / n
        if best is None or err < best[0]:
            best = (err, R, t)
    if best is None:
        return np.eye(3), np.zeros(3)
    return best[1], best[2]
