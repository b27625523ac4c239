"""Golden fixtures for the input-crop row (get_roi + ToTensor + Normalize), produced by the REFERENCE's own function
bodies: crop_square_resize / crop_resize / get_roi are taken from bop_dataset_pytorch.py with `ast` (the module header
imports imgaug, absent here) and run with the real cv2; the tensor step runs the reference's transform_pre recipe
(PIL -> torchvision ToTensor -> Normalize, bop_dataset_pytorch.py:334-347).  Outputs are large, so the fixture stores
their sha256 (the parity bar is bit-exact) plus three full crops for debugging.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_crop.py
"""
import ast
import hashlib
import os
import sys

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/zebrapose"
sys.path.insert(0, ROOT)

import cv2
import numpy as np
import torch
from PIL import Image
from torchvision import transforms

from workloads import synth_eval


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    ns = {"np": np, "cv2": cv2}
    tree = ast.parse(open(os.path.join(REF, "bop_dataset_pytorch.py")).read())
    names = ["crop_square_resize", "crop_resize", "get_roi", "padding_Bbox"]
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), "ref", "exec"), ns)
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
    img = synth_eval.make_image(5)
    boxes = synth_eval.make_crop_boxes(synth_eval.N_CROP_BOXES, 6)
    out = {"in_sha": np.array(sha(img, boxes))}
    for cs, method in synth_eval.CROP_CASES:
        u8_sha, f32_sha = [], []
        for i, b in enumerate(boxes):
            roi = ns["get_roi"](img, b, cs, interpolation=cv2.INTER_LINEAR, resize_method=method)
            t = tf(Image.fromarray(np.uint8(roi)).convert("RGB")).numpy()
            u8_sha.append(sha(roi)); f32_sha.append(sha(t))
            if cs == 64 and i < 3:
                out["full_u8_%d" % i] = roi
                out["full_f32_%d" % i] = t
        out["u8_%d_%s" % (cs, method)] = np.array(u8_sha)
        out["f32_%d_%s" % (cs, method)] = np.array(f32_sha)
    np.savez_compressed(os.path.join(HERE, "golden_crop_v1.npz"), **out)
    print("wrote golden_crop_v1.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
