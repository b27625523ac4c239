"""Golden fixture for class_base != 2 (CNN_output_to_pose.py:110 passes any base to class_code_images_to_class_id_image):
the REFERENCE's own CNN_outputs_to_object_pose on a base-3, 8-digit code image.  Runs only in the build container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_base3.py
"""
import os
import sys

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/zebrapose")

import numpy as np
from binary_code_helper.CNN_output_to_pose import (CNN_outputs_to_object_pose, build_non_unique_2D_3D_correspondence,
                                                   mapping_pixel_position_to_original_position)
from binary_code_helper.class_id_encoder_decoder import class_code_images_to_class_id_image


def make_case(seed=31, S=64, base=3, L=8):
    """a base-3 'dictionary' of 3^8 points on a sphere shell, a pose, and the digit image a perfect network would emit
    for it (nearest projected point per pixel), with 25 % of the masked pixels replaced by random codes"""
    import cv2
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    n = base ** L
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    pts = v * 45.0 * rng.uniform(0.7, 1.0, (n, 1))
    pts = np.array([[float("%.6g" % x) for x in row] for row in pts])
    pts[rng.random(n) < 0.1] = np.nan                                  # non-existing codes
    K = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1.0]])
    R = cv2.Rodrigues(rng.normal(size=3))[0]
    t = np.array([rng.uniform(-60, 60), rng.uniform(-40, 40), rng.uniform(500, 800)])
    ok = ~np.isnan(pts).any(1)
    P = (R @ pts[ok].T).T + t
    uv = (K @ P.T).T; uv = uv[:, :2] / uv[:, 2:]
    front = (P * ((R @ v[ok].T).T)).sum(1) < 0
    ids_ok = np.nonzero(ok)[0][front]
    x0, y0 = np.floor(uv.min(0)).astype(int) - 3; x1, y1 = np.ceil(uv.max(0)).astype(int) + 3
    side = int(max(x1 - x0, y1 - y0)); bbox = np.array([x0, y0, side, side])
    ys, xs = np.mgrid[0:S, 0:S]
    dist, nn = cKDTree(uv[front]).query(np.stack([(xs * side / S + x0).ravel(), (ys * side / S + y0).ravel()], 1))
    ids = ids_ok[nn].reshape(S, S)
    mask = (dist.reshape(S, S) < 1.5 * max(1, side / S)).astype(np.uint8)
    ids = np.where(rng.random((S, S)) < 0.25, rng.integers(0, n, (S, S)), ids)
    digits = np.stack([(ids // base ** (L - 1 - i)) % base for i in range(L)], -1).astype(np.float64)
    return pts, K, bbox, mask, digits, R, t


def main():
    base, L, S = 3, 8, 64
    pts, K, bbox, mask, digits, R, t = make_case(S=S, base=base, L=L)
    d = {float(i): pts[i].copy() for i in range(len(pts))}
    rot, tv, ok = CNN_outputs_to_object_pose(mask, digits, bbox, S, base, d, intrinsic_matrix=K)
    ids = class_code_images_to_class_id_image(digits, base)
    p2d, p3d = build_non_unique_2D_3D_correspondence(mask.nonzero(), ids, d)
    o2d = mapping_pixel_position_to_original_position(p2d, bbox, S)
    np.savez_compressed(os.path.join(HERE, "golden_base3_v1.npz"), pts=pts, K=K, bbox=bbox, mask=mask, digits=digits.astype(np.uint8),
                        R=np.asarray(rot), t=np.asarray(tv), ok=np.array(bool(ok)), uv=o2d.astype(np.float32), xyz=p3d.astype(np.float32),
                        R_gt=R, t_gt=t)
    print("base-3 fixture: %d correspondences, success %s" % (len(o2d), ok))


if __name__ == "__main__":
    main()
