"""CPU restatement of the decode half of the path (test infrastructure).

Every function cites the reference file:line it restates (paths relative to
/root/reference/zebrapose).  Two flavours are kept for the hot loop:
`build_correspondences_faithful` mirrors the reference's per-pixel Python dict
loop (this is what the reference costs on a CPU and what `bench.py` times as the
reference arm); `build_correspondences` is the vectorised equivalent the tests
use.  Both must agree bit-for-bit (tests/test_oracle_golden.py).
"""
import numpy as np


def load_dict_class_id_3D_points(path):
    """binary_code_helper/CNN_output_to_pose.py:10-32.  Header '<n_class> <base> <n_iter>',
    then '<id> <x> <y> <z>' lines; `line[:-1]` chops the last character of every line
    (so a final line without newline loses a digit -- reproduced)."""
    d = {}
    with open(path, "r") as f:
        a, b, c = f.readline().split(" ")
        total, base, n_iter = float(a), float(b), float(c)
        for line in f:
            line = line[:-1]
            code, x, y, z = line.split(" ")
            d[float(code)] = np.array([float(x), float(y), float(z)])
    return total, base, n_iter, d


def dict_to_table(d, n_bits):
    """dict{id -> xyz} -> float64 [2^n_bits, 3] table (row = class id).  Values may be (3,) or (1,3)."""
    n = 1 << n_bits
    tab = np.full((n, 3), np.nan, dtype=np.float64)
    for k, v in d.items():
        tab[int(k)] = np.asarray(v, dtype=np.float64).reshape(3)
    return tab


def table_to_dict(tab, float_keys=True):
    if float_keys:
        return {float(i): tab[i].copy() for i in range(len(tab))}
    return {int(i): tab[i].reshape(1, 3).copy() for i in range(len(tab))}


def generate_new_corres_table(tab, num_bit_old, num_bit_new):
    """binary_code_helper/generate_new_dict.py:4-33 on a table: parent = (zeros + sum of the 2^k
    children in ascending id order, float64) / 2^k ; NaN children propagate."""
    k = num_bit_old - num_bit_new
    kids = tab.reshape(1 << num_bit_new, 1 << k, 3)
    acc = np.zeros((1 << num_bit_new, 3))
    for j in range(1 << k):            # ascending id order, sequential float64 adds
        acc = acc + kids[:, j, :]
    return acc / (1 << k)


def generate_new_corres_dict(full_dict, num_bit_old, num_bit_new):
    """Same as the reference function: int keys, (1,3) values."""
    tab = generate_new_corres_table(dict_to_table(full_dict, num_bit_old), num_bit_old, num_bit_new)
    return table_to_dict(tab, float_keys=False)


def threshold_logits(logits):
    """common_ops.py:5-19 (BCE/L1 branch): sigmoid(x) > 0.5 -> 1.0 else 0.0 (float64 array).
    Product rule: float32(x) > 0; fixtures keep |x| >= 1e-6 so both agree."""
    x = np.asarray(logits, dtype=np.float32)
    return (x > 0).astype(np.float64)


def class_code_images_to_class_id_image(code_hwc, class_base=2):
    """binary_code_helper/class_id_encoder_decoder.py:17-28; channel 0 = MSB."""
    L = code_hwc.shape[2]
    out = np.zeros(code_hwc.shape[:2])
    for i in range(L):
        out = out + code_hwc[:, :, i] * (class_base ** (L - 1 - i))
    return out


def mapping_pixel_position_to_original_position(pixels, Bbox, Bbox_Size):
    """binary_code_helper/CNN_output_to_pose.py:34-50: float64 ratio*x + x0, astype(int) truncates."""
    rx = Bbox[2] / Bbox_Size
    ry = Bbox[3] / Bbox_Size
    ox = (rx * pixels[:, 0] + Bbox[0]).astype("int")
    oy = (ry * pixels[:, 1] + Bbox[1]).astype("int")
    return np.concatenate((ox.reshape(-1, 1), oy.reshape(-1, 1)), 1)


def build_correspondences_faithful(mask, class_id_image, d):
    """CNN_output_to_pose.py:53-64 + :111, same per-pixel Python loop / dict look-ups."""
    rows, cols = mask.nonzero()
    p2d = np.concatenate((cols.reshape(-1, 1), rows.reshape(-1, 1)), 1)
    ids = class_id_image[p2d[:, 1], p2d[:, 0]]
    p3d = np.zeros((p2d.shape[0], 3))
    for i in range(p2d.shape[0]):
        if np.isnan(np.array(d[ids[i]])).any():
            continue
        p3d[i] = np.array(d[ids[i]])
    return p2d, p3d


def build_correspondences(mask, class_id_image, tab):
    """Vectorised equivalent: row-major masked pixels, NaN rows -> (0,0,0), pixel kept."""
    rows, cols = mask.nonzero()
    p2d = np.stack([cols, rows], 1)
    ids = class_id_image[rows, cols].astype(np.int64)
    p3d = tab[ids].copy()
    p3d[np.isnan(p3d).any(1)] = 0.0
    return p2d, p3d


def decode_crop(mask, code_hwc, Bbox, Bbox_Size, tab):
    """CNN_output_to_pose.py:110-129 up to the float32 casts.  Returns (uv f32 [M,2], xyz f32 [M,3],
    class ids int64 [S,S])."""
    ids = class_code_images_to_class_id_image(code_hwc, 2)
    p2d, p3d = build_correspondences(mask, ids, tab)
    if len(p2d) == 0:
        return np.zeros((0, 2), np.float32), np.zeros((0, 3), np.float32), ids.astype(np.int64)
    o2d = mapping_pixel_position_to_original_position(p2d, Bbox, Bbox_Size)
    return o2d.astype(np.float32), p3d.astype(np.float32), ids.astype(np.int64)


# ---------------------------------------------------------------------------------------------
# north_star extension (NOT in the reference; parity unpinned): Hamming-nearest remap of
# non-existing codes.  Spec: remap[c] = existing code e minimising
# (popcount(c^e), c^e) lexicographically; existing codes map to themselves.
# ---------------------------------------------------------------------------------------------
def hamming_remap_table(exists):
    n = len(exists)
    nb = n.bit_length() - 1
    codes = np.arange(n, dtype=np.int64)
    remap = np.where(exists, codes, -1)
    todo = remap < 0
    if not exists.any():
        return np.zeros(n, np.uint16)
    xs = np.arange(1, n, dtype=np.int64)
    pc = np.zeros(n - 1, np.int64)
    for b in range(nb):
        pc += (xs >> b) & 1
    order = xs[np.lexsort((xs, pc))]            # xor patterns by (popcount, value)
    for x in order:
        if not todo.any():
            break
        cand = codes[todo] ^ x
        hit = exists[cand]
        idx = np.nonzero(todo)[0][hit]
        remap[idx] = cand[hit]
        todo[idx] = False
    return remap.astype(np.uint16)


def ignore_bit_table_hamming(tab16, k):
    """hamming mode for ignore_bit k: parent exists iff any child exists; value = float64 mean over the
    existing children (ascending id order); non-existing parents are then Hamming-remapped."""
    nb = int(np.log2(len(tab16)))
    kids = tab16.reshape(1 << (nb - k), 1 << k, 3)
    ex = ~np.isnan(kids).any(2)
    acc = np.zeros((kids.shape[0], 3))
    for j in range(1 << k):
        acc = acc + np.where(ex[:, j, None], kids[:, j, :], 0.0)
    cnt = ex.sum(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = acc / cnt[:, None]
    exists = cnt > 0
    remap = hamming_remap_table(exists)
    return mean[remap.astype(np.int64)], remap, exists
