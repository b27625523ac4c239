"""CPU restatement of the steps either side of the pose path (test infrastructure):
crop boxes (zebrapose/bop_dataset_pytorch.py:123-139, 162-194) and the BOP csv writer
(zebrapose/tools_for_BOP/write_to_cvs.py:6-62).  ADD / ADI live in oracle/metrics.py.  Pinned against the reference's
own functions by tests/golden/golden_eval_v1.npz (tests/golden/make_golden_eval.py)."""
import io
import math


def padding_box(box, ratio):
    """bop_dataset_pytorch.py:123-139 (float64 arithmetic, int() = truncation toward zero)"""
    x1, y1 = float(box[0]), float(box[1])
    x2, y2 = x1 + float(box[2]), y1 + float(box[3])
    cx, cy = 0.5 * (x1 + x2), 0.5 * (y1 + y2)
    pw, ph = math.trunc((x2 - x1) * ratio), math.trunc((y2 - y1) * ratio)
    return [math.trunc(cx - pw / 2), math.trunc(cy - ph / 2), pw, ph]


def final_box(box, method, max_x, max_y):
    """bop_dataset_pytorch.py:162-194"""
    x1, y1, bw, bh = (float(v) for v in box)
    x2, y2 = x1 + bw, y1 + bh
    if method in ("crop_square_resize", "crop_resize_by_warp_affine"):
        cx, cy = 0.5 * (x1 + x2), 0.5 * (y1 + y2)
        if bh > bw:
            x1, x2 = cx - bh / 2, cx + bh / 2
        else:
            y1, y2 = cy - bw / 2, cy + bw / 2
    elif method == "crop_resize":
        x1, y1, x2, y2 = max(x1, 0), max(y1, 0), min(x2, max_x), min(y2, max_y)
    else:
        return list(box)
    x1, y1, x2, y2 = (math.trunc(v) for v in (x1, y1, x2, y2))
    return [x1, y1, x2 - x1, y2 - y1]


def bop_csv_text(obj_id, scene_ids, img_ids, Rs, ts, scores):
    """write_to_cvs.py:6-62 as a string"""
    f = io.StringIO()
    f.write("scene_id,im_id,obj_id,score,R,t,time\n")
    for s, i, r, t, sc in zip(scene_ids, img_ids, Rs, ts, scores):
        if sc == -1:
            continue
        f.write(",".join([str(s), str(i), str(obj_id), str(sc)]) + ",")
        f.write(" ".join(str(r[a][b]) for a in range(3) for b in range(3)) + ",")
        f.write(" ".join(str(t[a][0]) for a in range(3)) + ",-1\n")
    return f.getvalue()


# ---- input crops: get_roi + ToTensor + Normalize ------------------------------------------
import numpy as np


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 images, restated from OpenCV's fixed-point
    path (imgproc/resize.cpp: float32 tap positions, 11-bit coefficients rounded half-to-even, columns clamped with the
    fraction reset, rows clamped WITHOUT resetting the fraction, vertical pass
    (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2; an exact 2x2 decimation switches to INTER_AREA's rounded mean).
    Pinned bit-exactly against cv2 4.13 in tests/test_oracle_evalside.py."""
    sh, sw = src.shape[:2]
    cn = 1 if src.ndim == 2 else src.shape[2]
    s = src.reshape(sh, sw, cn).astype(np.int64)
    shape = (dh, dw) if src.ndim == 2 else (dh, dw, cn)
    if sw == 2 * dw and sh == 2 * dh:
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
        return out.astype(np.uint8).reshape(shape)
    scale_x, scale_y = 1.0 / (dw / sw), 1.0 / (dh / sh)
    d = np.arange(dw)
    f = ((d + 0.5) * scale_x - 0.5).astype(np.float32)
    sx = np.floor(f).astype(np.int64)
    f = (f - sx.astype(np.float32)).astype(np.float32)
    lo = sx < 0
    f[lo] = 0; sx[lo] = 0
    hi = sx >= sw - 1
    f[hi] = 0; sx[hi] = sw - 1
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    sx1 = np.minimum(sx + 1, sw - 1)
    d = np.arange(dh)
    f = ((d + 0.5) * scale_y - 0.5).astype(np.float32)
    sy = np.floor(f).astype(np.int64)
    f = (f - sy.astype(np.float32)).astype(np.float32)
    b0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    b1 = np.rint(f * np.float32(2048)).astype(np.int64)
    r0, r1 = np.clip(sy, 0, sh - 1), np.clip(sy + 1, 0, sh - 1)

    def hpass(rows):
        return rows[:, sx, :] * a0[None, :, None] + rows[:, sx1, :] * a1[None, :, None]

    S0, S1 = hpass(s[r0]), hpass(s[r1])
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8).reshape(shape)


def crop_square(img, box):
    """the zero-padded square canvas of crop_square_resize (bop_dataset_pytorch.py:36-70), before the resize"""
    x1, y1 = int(box[0]), int(box[1])
    bw, bh = max(int(box[2]), 0), max(int(box[3]), 0)
    fx1, fx2, fy1, fy2 = float(x1), float(x1 + bw), float(y1), float(y1 + bh)
    cx, cy = 0.5 * (fx1 + fx2), 0.5 * (fy1 + fy2)
    if bh > bw:
        fx1, fx2 = cx - bh / 2, cx + bh / 2
    else:
        fy1, fy2 = cy - bw / 2, cy + bw / 2
    x1, y1, x2, y2 = math.trunc(fx1), math.trunc(fy1), math.trunc(fx2), math.trunc(fy2)
    side = max(bh, bw)
    roi = np.zeros((side, side) + img.shape[2:], img.dtype)
    rx1 = max(-x1, 0); x1 = max(x1, 0)
    rx2 = rx1 + min(img.shape[1] - x1, x2 - x1)
    ry1 = max(-y1, 0); y1 = max(y1, 0)
    ry2 = ry1 + min(img.shape[0] - y1, y2 - y1)
    x2, y2 = min(x2, img.shape[1]), min(y2, img.shape[0])
    roi[ry1:ry2, rx1:rx2] = img[y1:y2, x1:x2]
    return roi


def get_roi_u8(img, box, crop_size, method):
    """get_roi(..., interpolation=cv2.INTER_LINEAR, resize_method) for uint8 images (bop_dataset_pytorch.py:110-121)"""
    if method == "crop_square_resize":
        return resize_linear_u8(crop_square(img, box), crop_size, crop_size)
    if method == "crop_resize":
        x1, x2 = max(0, int(box[0])), min(img.shape[1], int(box[0]) + int(box[2]))
        y1, y2 = max(0, int(box[1])), min(img.shape[0], int(box[1]) + int(box[3]))
        return resize_linear_u8(img[y1:y2, x1:x2], crop_size, crop_size)
    raise NotImplementedError(method)


MEAN = np.array([0.485, 0.456, 0.406], np.float32)
STD = np.array([0.229, 0.224, 0.225], np.float32)


def to_tensor_normalize(roi_u8):
    """transforms.ToTensor() + Normalize(mean, std) (bop_dataset_pytorch.py:334-347): float32 CHW"""
    x = roi_u8.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    return (x - MEAN[:, None, None]) / STD[:, None, None]
