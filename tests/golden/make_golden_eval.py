"""Golden fixtures for the steps either side of the pose path (ADD / ADI, crop boxes, BOP csv), produced by the
REFERENCE's own function bodies.  lib/pysixd/pose_error.py and bop_dataset_pytorch.py cannot be imported here (their
module headers pull mmcv / imgaug / termcolor, absent from the image), so the function definitions are taken from the
reference files with `ast` at generation time and executed unmodified in a namespace that holds only numpy and
scipy.spatial -- nothing of the reference is copied into the repo.  tools_for_BOP/write_to_cvs.py imports cleanly.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_eval.py
"""
import ast
import hashlib
import os
import sys
import tempfile

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/zebrapose"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import numpy as np
from scipy import spatial

from workloads import synth_eval


def ref_functions(path, names, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    ns = {"np": np, "spatial": spatial}
    ref_functions(os.path.join(REF, "lib/pysixd/misc.py"), ["transform_pts_Rt"], ns)
    ref_functions(os.path.join(REF, "lib/pysixd/pose_error.py"), ["add", "adi"], ns)
    ref_functions(os.path.join(REF, "bop_dataset_pytorch.py"), ["padding_Bbox", "get_final_Bbox"], ns)
    from tools_for_BOP import write_to_cvs
    out = {}

    # --- ADD / ADI ---------------------------------------------------------------------------------------------
    for tag, V, seed in synth_eval.MODELS:
        pts = synth_eval.make_model(V, seed)
        est, gt = synth_eval.make_pose_pairs(synth_eval.N_PAIRS, seed + 1)
        add = np.array([ns["add"](e[:9].reshape(3, 3), e[9:].reshape(3, 1), g[:9].reshape(3, 3), g[9:].reshape(3, 1), pts)
                        for e, g in zip(est, gt)])
        adi = np.array([ns["adi"](e[:9].reshape(3, 3), e[9:].reshape(3, 1), g[:9].reshape(3, 3), g[9:].reshape(3, 1), pts)
                        for e, g in zip(est, gt)])
        out["err_%s_add" % tag] = add
        out["err_%s_adi" % tag] = adi
        out["err_%s_sha" % tag] = np.array(sha(pts, est, gt))

    # --- crop boxes --------------------------------------------------------------------------------------------
    boxes = synth_eval.make_boxes(synth_eval.N_BOXES, 7)
    out["box_sha"] = np.array(sha(boxes))
    for ratio in synth_eval.PAD_RATIOS:
        out["box_pad_%g" % ratio] = np.array([ns["padding_Bbox"](b if i % 2 else b.astype(np.int64), ratio)
                                              for i, b in enumerate(boxes)], np.int64)
        for method in synth_eval.METHODS:
            out["box_final_%g_%s" % (ratio, method)] = np.array(
                [ns["get_final_Bbox"](ns["padding_Bbox"](b if i % 2 else b.astype(np.int64), ratio), method, 640, 480)
                 for i, b in enumerate(boxes)], np.int64)
    for method in synth_eval.METHODS:      # get_final_Bbox alone, on raw (float and int) boxes
        out["box_finalonly_%s" % method] = np.array(
            [ns["get_final_Bbox"](b if i % 2 else b.astype(np.int64), method, 640, 480) for i, b in enumerate(boxes)], np.float64)

    # --- BOP csv -----------------------------------------------------------------------------------------------
    scene, img, Rs, ts, scores = synth_eval.make_csv_rows(9)
    with tempfile.TemporaryDirectory() as d:
        write_to_cvs.write_cvs(d, "lmo_ape", 1, scene, img, Rs, ts, scores)
        out["csv_text"] = np.array(open(os.path.join(d, "lmo_ape.csv")).read())

    np.savez_compressed(os.path.join(HERE, "golden_eval_v1.npz"), **out)
    print("wrote golden_eval_v1.npz:", {k: (v.shape if v.ndim else str(v)[:16]) for k, v in out.items()})


if __name__ == "__main__":
    main()
