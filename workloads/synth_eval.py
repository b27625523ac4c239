"""Seeded synthetic inputs for the evaluation-side fixtures (test infrastructure): model vertices, pose pairs,
detection boxes, csv rows.  Shared by tests/golden/make_golden_eval.py and the tests so both see identical inputs."""
import numpy as np
import cv2

MODELS = [("v1500", 1500, 21), ("v5841", 5841, 22)]     # (tag, vertices, seed); 5841 = LM-O ape model size
N_PAIRS = 6
N_BOXES = 120
PAD_RATIOS = (1.5, 1.2)
METHODS = ("crop_resize", "crop_square_resize", "crop_resize_by_warp_affine")


def make_model(V, seed):
    """vertices of a bumpy ellipsoid, ~100 mm across, 6-decimal values like a .ply file"""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(V, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = 1.0 + 0.15 * np.sin(5 * d[:, :1]) * np.cos(3 * d[:, 1:2])
    return np.round(d * r * np.array([55.0, 40.0, 30.0]), 6)


def make_pose_pairs(n, seed):
    """(est, gt) float64 [n,12]: gt random; est = gt perturbed by 0.05 .. 20 degrees / 0.1 .. 30 mm (one pair identical)"""
    rng = np.random.default_rng(seed)
    est, gt = [], []
    for i in range(n):
        Rg = cv2.Rodrigues(rng.normal(size=3))[0]
        tg = np.array([rng.uniform(-100, 100), rng.uniform(-80, 80), rng.uniform(600, 1200)])
        ang = np.radians([0.0, 0.05, 0.5, 2.0, 8.0, 20.0][i % 6])
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        Re = cv2.Rodrigues(ax * ang)[0] @ Rg
        te = tg + rng.normal(size=3) * [0.0, 0.1, 0.5, 2.0, 8.0, 30.0][i % 6]
        gt.append(np.concatenate([Rg.ravel(), tg]))
        est.append(np.concatenate([Re.ravel(), te]))
    return np.array(est), np.array(gt)


def make_boxes(n, seed):
    """detection boxes x,y,w,h: partly outside a 640x480 image, negative origins, fractional values on odd rows"""
    rng = np.random.default_rng(seed)
    b = np.stack([rng.uniform(-60, 600, n), rng.uniform(-60, 440, n), rng.uniform(3, 300, n), rng.uniform(3, 300, n)], 1)
    b[::2] = np.round(b[::2])
    b[5] = [10, 20, 50, 50]          # square
    b[6] = [-30, -40, 700, 500]      # larger than the image
    return b


def make_csv_rows(seed):
    rng = np.random.default_rng(seed)
    n = 7
    Rs = [cv2.Rodrigues(rng.normal(size=3))[0] for _ in range(n)]
    Rs[2] = np.eye(3)                                   # failed crop: identity / zeros (test.py:457-463)
    ts = [rng.normal(size=(3, 1)) * [[100], [80], [900]] for _ in range(n)]
    ts[2] = np.zeros((3, 1))
    scene = [int(v) for v in rng.integers(1, 60, n)]
    img = [int(v) for v in rng.integers(0, 2000, n)]
    scores = [1, 0.75, 1, -1, np.float64(0.5), 1, np.float32(0.25)]
    return scene, img, Rs, ts, scores


MERGE_FILES = [("lmo", "ape", 1, 9), ("lmo", "can", 5, 10), ("ycbv", "002_master_chef_can", 1, 11)]   # (dataset, object, obj_id, rows seed)


CROP_CASES = [(256, "crop_square_resize"), (256, "crop_resize"), (128, "crop_square_resize"), (64, "crop_resize")]
N_CROP_BOXES = 14


def make_image(seed, H=480, W=640):
    """uint8 RGB test image: smooth gradients + texture + noise (edges and flat areas both present)"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W]
    img = np.stack([127 + 120 * np.sin(x / 37.0) * np.cos(y / 23.0), (x * 255.0 / W + y) % 256, 40 + 0.3 * ((x // 16 + y // 16) % 2) * 600], -1)
    img = img + rng.normal(0, 12, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def make_crop_boxes(n, seed, H=480, W=640):
    """integer boxes x,y,w,h as padding_Bbox returns them: inside, partly outside on every side, tall, wide, tiny,
    exactly 2 x the crop size (the INTER_AREA switch of cv2.resize)"""
    rng = np.random.default_rng(seed)
    b = np.stack([rng.integers(-80, W - 40, n), rng.integers(-80, H - 40, n), rng.integers(8, 330, n), rng.integers(8, 330, n)], 1)
    b[0] = [100, 60, 256, 256]; b[1] = [40, 30, 512, 512]; b[2] = [-50, -40, 200, 120]; b[3] = [500, 380, 300, 90]
    b[4] = [320, 200, 9, 31]; b[5] = [10, 10, 128, 128]
    return b.astype(np.int64)
