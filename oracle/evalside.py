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
