"""Mirror of zebrapose/binary_code_helper/class_id_encoder_decoder.py (decode side: :17-28, :61-87)."""
import numpy as np


def class_code_images_to_class_id_image(class_code_images, class_base=2):
    """HWC code image (values 0..base-1, channel 0 most significant) -> float64 [H,W] class ids
    (class_id_encoder_decoder.py:17-28).  Computed on the device (zp_codes_to_ids)."""
    from zebrapose_b200.engine import default_engine
    a = np.asarray(class_code_images, dtype=np.float64)
    H, W, L = a.shape
    ids = default_engine().codes_to_ids(a.reshape(H * W, L), class_base)
    return ids.cpu().numpy().reshape(H, W)


def code_to_id(class_code, class_base=2):
    """class_id_encoder_decoder.py:65-75"""
    v = 0
    n = len(class_code)
    for i, c in enumerate(class_code):
        v = v + c * (class_base ** (n - 1 - i))
    return v


def str_code_to_id(str_class_code, class_base=2):
    """class_id_encoder_decoder.py:77-87"""
    return code_to_id([int(c) for c in str_class_code], class_base)
