"""The helpers of the reference's dataset module that sit right before the network and produce the `Bbox` argument of
the pose path (zebrapose/bop_dataset_pytorch.py:123-139 padding_Bbox, :162-194 get_final_Bbox, :110-121 get_roi; called
per detection at test_vivo.py:147-159), with the reference names and argument order.  They run the device kernels behind
zp_final_bbox / zp_crop_input; the batched forms that keep everything on the GPU are Engine.final_bboxes() and
Engine.crop_inputs() (which also applies ToTensor + Normalize)."""
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


INTER_LINEAR = 1          # cv2.INTER_LINEAR


def get_roi(input, Bbox, crop_size, interpolation, resize_method):
    """bop_dataset_pytorch.py:110-121 for the network input: uint8 [H,W,3] image, cv2.INTER_LINEAR,
    resize_method "crop_resize" | "crop_square_resize" -> uint8 [crop_size,crop_size,3], bit-identical to cv2's."""
    img = np.asarray(input)
    if interpolation != INTER_LINEAR or img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise NotImplementedError("the device path crops uint8 RGB images with cv2.INTER_LINEAR (the network input); "
                                  "GT / mask crops (INTER_NEAREST) are training-side")
    _, u8 = default_engine().crop_inputs(img, np.asarray(Bbox, np.float64).reshape(1, 4), crop_size=crop_size,
                                         resize_method=resize_method, return_u8=True)
    return u8[0].cpu().numpy()
