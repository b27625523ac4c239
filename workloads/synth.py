"""Synthetic inputs for the parity tests and the bench (test infrastructure).

Everything is seeded with numpy.random.default_rng so CPU and GPU arms see identical inputs.
No file under /root/reference is read here.
"""
import numpy as np
import cv2
from scipy.spatial import cKDTree

LM_K = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1.0]])
YCBV_K = np.array([[1066.778, 0, 312.9869], [0, 1067.487, 241.3109], [0, 0, 1.0]])
TLESS_K = np.array([[1075.65091572, 0, 360.0], [0, 1073.90347929, 270.0], [0, 0, 1.0]])


def make_dict(n_bits=16, seed=0, radius=50.0, missing_frac=0.0):
    """Random 'mesh' dictionary: 2^n_bits points on a bumpy sphere shell, 6-significant-digit values as the
    generator writes them (Generate_Mesh_with_GT_Color.cpp:616-624), optional NaN (non-existing) rows."""
    rng = np.random.default_rng(seed)
    n = 1 << n_bits
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    pts = v * radius * rng.uniform(0.7, 1.0, (n, 1))
    # hierarchical codes like Generate_Mesh_with_GT_Color.cpp's balanced 2-means (:61, :396): bit l splits every
    # group of the previous level in two equal halves along its widest axis, so sibling codes are spatial neighbours
    # and the ignore-bit parents (mean of the children) stay on the surface
    order = np.arange(n)
    for level in range(n_bits):
        g = pts[order].reshape(1 << level, n >> level, 3)
        axis = g.var(1).argmax(1)
        key = np.take_along_axis(g, axis[:, None, None], 2)[:, :, 0]
        sub = np.argsort(key, 1, kind="stable")
        order = np.take_along_axis(order.reshape(1 << level, n >> level), sub, 1).ravel()
    pts, v = pts[order], v[order]
    pts = np.array([[float("%.6g" % c) for c in p] for p in pts]) if n <= 4096 else _round6(pts)
    missing = rng.random(n) < missing_frac
    tab = pts.copy()
    tab[missing] = np.nan
    return tab, v, missing


def _round6(a):
    """vectorised '%.6g' round trip"""
    with np.errstate(divide="ignore"):
        mag = np.where(a == 0, 0, np.floor(np.log10(np.abs(a))))
    scale = 10.0 ** (5 - mag)
    r = np.round(a * scale) / scale
    # make sure the value equals the decimal text round trip
    return np.array([float(s) for s in np.char.mod("%.6g", r.ravel())]).reshape(a.shape)


def write_dict_file(path, tab, n_bits=16, final_newline=True):
    """Class_CorresPoint%06d.txt as written by Generate_Mesh_with_GT_Color.cpp:616-624."""
    lines = ["%d 2 %d" % (len(tab), n_bits)]
    for i, p in enumerate(tab):
        if np.isnan(p).any():
            lines.append("%d nan -nan nan" % i)
        else:
            lines.append("%d %.6g %.6g %.6g" % (i, p[0], p[1], p[2]))
    txt = "\n".join(lines) + ("\n" if final_newline else "")
    with open(path, "w") as f:
        f.write(txt)


def random_pose(rng, zmin=600.0, zmax=1200.0):
    R = cv2.Rodrigues(rng.normal(size=3))[0]
    t = np.array([rng.uniform(-100, 100), rng.uniform(-80, 80), rng.uniform(zmin, zmax)])
    return R, t


def make_crop(tab, nrm, seed, S=128, n_bits=16, outlier=0.3, K=LM_K, bitflip=0.02, R=None, t=None):
    """One synthetic crop: GT pose, nearest-front-facing-vertex code image, mask, outliers, bit flips.
    Returns dict(mask u8[S,S], bits u8[S,S,n_bits] (MSB first), bbox int64[4] (x,y,w,h), K, R, t)."""
    rng = np.random.default_rng(seed)
    pts = np.where(np.isnan(tab), 0.0, tab)
    valid = ~np.isnan(tab).any(1)
    if R is None:
        R, t = random_pose(rng)
    P = (R @ pts.T).T + t
    uv = (K @ P.T).T
    uv = uv[:, :2] / uv[:, 2:]
    x0, y0 = np.floor(uv[valid].min(0)).astype(int) - 3
    x1, y1 = np.ceil(uv[valid].max(0)).astype(int) + 3
    side = int(max(x1 - x0, y1 - y0))
    bbox = np.array([x0, y0, side, side], dtype=np.int64)
    front = ((P * ((R @ nrm.T).T)).sum(1) < 0) & valid
    idxf = np.nonzero(front)[0]
    tree = cKDTree(uv[front])
    ys, xs = np.mgrid[0:S, 0:S]
    q = np.stack([(xs * side / S + x0).ravel(), (ys * side / S + y0).ravel()], 1)
    dist, nn = tree.query(q)
    ids = idxf[nn].reshape(S, S)
    mask = (dist.reshape(S, S) < 1.5 * max(1.0, side / S)).astype(np.uint8)
    n = len(tab)
    ids = np.where(rng.random((S, S)) < outlier, rng.integers(0, n, (S, S)), ids)
    bits = ((ids[..., None] >> (n_bits - 1 - np.arange(n_bits))) & 1).astype(np.uint8)
    if bitflip:
        bits ^= (rng.random(bits.shape) < bitflip).astype(np.uint8)
    return dict(mask=mask, bits=bits, bbox=bbox, K=K.copy(), R=R, t=t, seed=seed)


def crop_to_logits(crop, seed=None, lo=1.0, hi=6.0, dtype=np.float32):
    """[1+n_bits, S, S] logits: channel 0 = mask, 1.. = bits MSB first; sign from the bit, |x| in [lo,hi)
    (|x| >= 1e-6 rule)."""
    rng = np.random.default_rng(crop["seed"] * 7919 + 13 if seed is None else seed)
    planes = np.concatenate([crop["mask"][None], crop["bits"].transpose(2, 0, 1)], 0).astype(bool)
    mag = rng.uniform(lo, hi, planes.shape)
    return np.where(planes, mag, -mag).astype(dtype)


def make_batch(B, S=128, n_bits=16, n_dicts=1, seed=1000, K=LM_K, outlier=0.3, bitflip=0.02,
               missing_frac=0.0, radius=(51.0, 51.0)):
    """A batch of crops over `n_dicts` dictionaries.  Returns (logits f32 [B,1+n_bits,S,S], bboxes i64 [B,4],
    Ks f64 [B,3,3], obj_ids i32 [B], tables list of f64 [2^n_bits,3], crops list)."""
    rng = np.random.default_rng(seed)
    dicts = []
    for j in range(n_dicts):
        r = rng.uniform(radius[0], radius[1])
        dicts.append(make_dict(n_bits, seed * 131 + j, radius=r, missing_frac=missing_frac))
    obj = rng.integers(0, n_dicts, B).astype(np.int32)
    logits = np.empty((B, 1 + n_bits, S, S), np.float32)
    bboxes = np.empty((B, 4), np.int64)
    Ks = np.empty((B, 3, 3))
    crops = []
    for i in range(B):
        tab, nrm, _ = dicts[obj[i]]
        c = make_crop(tab, nrm, seed * 65536 + i, S=S, n_bits=n_bits, outlier=outlier, K=K, bitflip=bitflip)
        crops.append(c)
        logits[i] = crop_to_logits(c)
        bboxes[i] = c["bbox"]
        Ks[i] = c["K"]
    return logits, bboxes, Ks, obj, [d[0] for d in dicts], crops
