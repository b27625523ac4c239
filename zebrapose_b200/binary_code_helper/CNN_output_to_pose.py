"""Mirror of zebrapose/binary_code_helper/CNN_output_to_pose.py: same names, argument order and return conventions;
the per-pixel dictionary loop (:53-64) and cv2.solvePnPRansac (:155-158) run on the B200 through
libzebrapose_b200.so.  One crop per call like the reference; the batched device entry
(zebrapose_b200.Engine.decode_and_pose_batch) is what a throughput-minded caller switches to."""
import numpy as np
import torch

from zebrapose_b200.engine import default_engine, dict_to_table

USE_PYPROGRESSIVEX = False   # the Progressive-X branch (:133-152) is not provided; RANSAC-EPnP is the cv2 branch

_DICT_SLOT = 255             # table slot the per-crop drop-in uses
_dict_cache = {}             # device index -> (key, dictionary object kept alive)


def load_dict_class_id_3D_points(path):
    """CNN_output_to_pose.py:10-32.  -> (n_class, base, n_iter as floats, dict{float id: float64[3]}).  The reference
    drops the last character of every line; a final line without a newline therefore loses a digit (kept)."""
    with open(path, "r") as f:
        text = f.read()
    header, _, body = text.partition("\n")
    n_class, base, n_iter = (float(v) for v in header.split(" "))
    table = {}
    rows = body.split("\n")
    complete = rows[:-1]                 # every row that ended in a newline
    tail = rows[-1][:-1]                 # unterminated last row: last character chopped
    for row in complete + ([tail] if rows[-1] != "" else []):
        code, x, y, z = row.split(" ")
        table[float(code)] = np.array([float(x), float(y), float(z)])
    return n_class, base, n_iter, table


def mapping_pixel_position_to_original_position(pixels, Bbox, Bbox_Size):
    """CNN_output_to_pose.py:34-50: crop pixel (x,y) -> original image pixel, float64 `w/S*x + x0`, truncation.
    Device kernel zp_remap_pixels; returns an int64 numpy array [N,2]."""
    px = np.asarray(pixels)
    if px.shape[0] == 0:
        return np.zeros((0, 2), dtype=np.int64)
    bb = np.asarray(Bbox.cpu() if isinstance(Bbox, torch.Tensor) else Bbox, dtype=np.float64)
    return default_engine().remap_pixels(px[:, :2], bb, Bbox_Size).cpu().numpy()


def _engine_with_dict(d, n_bits):
    eng = default_engine()               # one engine per device; the upload cache is per device too
    key = (id(d), len(d), n_bits)
    if _dict_cache.get(eng.device.index, (None,))[0] != key:
        eng.upload_dict(_DICT_SLOT, dict_to_table(d, n_bits), n_bits=n_bits, ignore_bit=0, nonexist="zero")
        _dict_cache[eng.device.index] = (key, d)      # keep the object alive so id() stays unique
    return eng


def CNN_outputs_to_object_pose(mask_image, class_code_image, Bbox, Bbox_Size, class_base=2,
                               dict_class_id_3D_points=None, intrinsic_matrix=None, return_info=False):
    """CNN_output_to_pose.py:100-160.  mask_image [S,S] (uint8 | float, != 0 is foreground, may be an external
    mask), class_code_image [S,S,L] of 0/1 (already sliced to 16-k channels when ignore_bit = k, with the matching
    dictionary), Bbox = [x,y,w,h], intrinsic_matrix 3x3 (numpy or torch; default LM intrinsics).
    -> (rot 3x3 float64, tvecs 3x1 float64 [mm], success) or ([], [], False) when fewer than 6 correspondences.

    How close to the reference's cv2.solvePnPRansac (:155-157), per crop: the RANSAC hypotheses are bit-identical to cv2's
    (the minimal solver replays OpenCV's EPnP arithmetic), near-ties of cv2's strictly-greater record rule are decided on
    counts taken with cv2's own arithmetic, so the winner, the iteration count and the final inlier set are cv2's, and the
    pose agrees within 0.05 deg / 0.5 mm on all 1536 measured crops at ignore_bit 0 / 2 / 4 (profiles/r2x_parity_*.json;
    largest difference 2.4e-6 deg / 1.2e-8 mm: the final EPnP on the inliers sums in another order than cv2).  Pinned per
    crop by tests/test_gpu_ransac.py (winner / iterations / inliers equal cv2's) and tests/test_gpu_dropin.py (tolerance on
    every golden crop)."""
    if intrinsic_matrix is None:
        intrinsic_matrix = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1.0]])
    K = np.asarray(intrinsic_matrix.cpu() if isinstance(intrinsic_matrix, torch.Tensor) else intrinsic_matrix, np.float64)
    bb = np.asarray(Bbox.cpu() if isinstance(Bbox, torch.Tensor) else Bbox, np.float64).reshape(1, 4)
    code = np.asarray(class_code_image)
    S, _, L = code.shape
    class_base = int(class_base)
    planes = torch.from_numpy(np.ascontiguousarray(code.transpose(2, 0, 1)))
    if class_base == 2:
        eng = _engine_with_dict(dict_class_id_3D_points, L)
        dev = eng.device
        # thresholded host arrays -> +-1 "logits" so the same decode kernel applies (x > 0)
        logits = torch.where(planes.to(dev) != 0, 1.0, -1.0).to(torch.float32).unsqueeze(0).contiguous()
        mask = torch.from_numpy(np.ascontiguousarray(np.asarray(mask_image) != 0)).to(dev).reshape(1, S, S)
        corr, counts = eng.decode(logits, bb, None, obj_default=_DICT_SLOT, mask_ch=0, bit0_ch=0, n_bits=L, ignore_bit=0,
                                  ext_mask=mask)
    else:
        # any base (the CE ablation heads, :110 -> class_id_encoder_decoder.py:17-28): digit d of a pixel becomes a one-hot
        # group of `base` logits, which the CE decode kernel (first maximum of the group) maps back to d
        n_classes = class_base ** L
        if n_classes > 65536:
            raise ValueError("class_base ** code length = %d exceeds the 16-bit class ids of the device path" % n_classes)
        n_bits = max(1, int(np.ceil(np.log2(n_classes))))
        eng = _engine_with_dict(dict_class_id_3D_points, n_bits)
        dev = eng.device
        digits = planes.to(dev).round().long().clamp(0, class_base - 1)                       # [L,S,S]
        onehot = torch.nn.functional.one_hot(digits, class_base).permute(0, 3, 1, 2)          # [L,base,S,S]
        logits = (onehot.reshape(1, L * class_base, S, S).to(torch.float32) * 2 - 1).contiguous()
        mask = torch.from_numpy(np.ascontiguousarray(np.asarray(mask_image) != 0)).to(dev).reshape(1, S, S)
        corr, counts = eng.decode_ce(logits, bb, None, base=class_base, n_digits=L, obj_default=_DICT_SLOT, mask_ch=0,
                                     digit0_ch=0, ext_mask=mask)
    r = eng.ransac(corr, counts, K.reshape(1, 9), H=150, m=5, thr=2.0, conf=0.99, sampler="cv2",
                   select="cv2_replay", final="epnp", return_details=return_info)
    n = int(counts.item())
    if n < 6:                                        # :119,126 -> rot = [], tvecs = [], success = False
        return ([], [], False, {}) if return_info else ([], [], False)
    pose = r["poses"][0].cpu().numpy()
    rot = pose[:9].reshape(3, 3).copy()
    tvecs = pose[9:].reshape(3, 1).copy()
    if return_info:
        info = dict(n_correspondences=n, n_inliers=int(r["n_inliers"].item()), status=int(r["status"].item()),
                    coord_2d=corr[0, 0:2, :n].T.cpu().numpy(), coord_3d=corr[0, 2:5, :n].T.cpu().numpy(),
                    inlier_mask=r["inlier_mask"][0, :n].cpu().numpy().astype(bool))
        return rot, tvecs, True, info
    return rot, tvecs, True                          # success stays True even if RANSAC found no model (:155)


def CNN_outputs_to_object_info(*args, **kwargs):
    """north_star name for the same function; returns (rot, tvecs, success, info)."""
    kwargs["return_info"] = True
    return CNN_outputs_to_object_pose(*args, **kwargs)
