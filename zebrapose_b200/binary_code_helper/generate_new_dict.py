"""Mirror of zebrapose/binary_code_helper/generate_new_dict.py:4-33 (one-off host work; the device tables are built
from the FULL dictionary by Engine.upload_dict(ignore_bit=k), so the batched path never needs this dict)."""
import numpy as np


def generate_new_corres_dict(full_binary_corres_dict, num_bit_old_dict, num_bit_new_dict):
    """-> dict{int prefix id: float64 (1,3)}: mean of the 2^k children (zeros(1,3) + children in ascending id order,
    float64, / 2^k); a NaN child makes the parent NaN."""
    k = num_bit_old_dict - num_bit_new_dict
    n_child = 1 << k
    out = {}
    for prefix in range(1 << num_bit_new_dict):
        acc = np.zeros((1, 3))
        for child in range(prefix * n_child, (prefix + 1) * n_child):
            acc = acc + full_binary_corres_dict[child]
        out[prefix] = acc / n_child
    return out
