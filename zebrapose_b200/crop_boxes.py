"""The two box helpers of the reference's dataset module that produce the `Bbox` argument of the pose path
(zebrapose/bop_dataset_pytorch.py:123-139 padding_Bbox, :162-194 get_final_Bbox; called per detection at
test_vivo.py:147-150), with the reference names and argument order.  Both run the device kernel behind zp_final_bbox;
the batched form that keeps the boxes on the GPU is Engine.final_bboxes()."""
import numpy as np

from .engine import default_engine


def padding_Bbox(Bbox, padding_ratio):
    """bop_dataset_pytorch.py:123-139 -> np.array([x, y, w, h]) of Python ints"""
    out = default_engine().final_bboxes(np.asarray(Bbox, np.float64).reshape(1, 4), padding_ratio, "none")
    return out[0].cpu().numpy().astype(np.int64)


def get_final_Bbox(Bbox, resize_method, max_x, max_y):
    """bop_dataset_pytorch.py:162-194.  An unknown resize_method returns the box unchanged, like the reference."""
    if resize_method not in ("crop_resize", "crop_square_resize", "crop_resize_by_warp_affine"):
        return Bbox
    out = default_engine().final_bboxes(np.asarray(Bbox, np.float64).reshape(1, 4), 0.0, resize_method, max_x, max_y)
    return out[0].cpu().numpy().astype(np.int64)
