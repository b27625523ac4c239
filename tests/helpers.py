"""Shared helpers for the parity tests."""
import hashlib

import numpy as np

from workloads import synth

GOLDEN_CROPS = {  # tag -> (table key, seed, S, ignore_bit)
    "c1_full": ("full", 1001 * 65536 + 0, 128, 0),
    "c1_nan20": ("nan20", 1001 * 65536 + 1, 128, 0),
    "c3_k1": ("nan20", 1003 * 65536 + 0, 128, 1),
    "c3_k4": ("nan20", 1003 * 65536 + 1, 128, 4),
    "c3_k8": ("nan20", 1003 * 65536 + 2, 128, 8),
    "s64_k0": ("full", 77, 64, 0),
    "c3f_k2": ("full", 1003 * 65536 + 10, 128, 2),
    "c3f_k4": ("full", 1003 * 65536 + 11, 128, 4),
}


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def regen_crop(tables, tag):
    key, seed, S, k = GOLDEN_CROPS[tag]
    tab, nrm = tables[key]
    c = synth.make_crop(tab, nrm, seed, S=S)
    logits = synth.crop_to_logits(c)
    return tab, c, logits, S, k


def as_set(uv, xyz):
    """order-independent multiset view of a correspondence list (bit patterns)"""
    a = np.concatenate([np.asarray(uv, np.float32), np.asarray(xyz, np.float32)], 1)
    a = np.ascontiguousarray(a).view(np.uint32).reshape(len(a), 5)
    return a[np.lexsort(a.T[::-1])]
