"""CPU: the evaluation-side restatements (oracle/metrics.py ADD / ADI, oracle/evalside.py boxes + csv) and the pure-Python
csv mirror against the fixtures produced by the reference's own functions (tests/golden/make_golden_eval.py)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import evalside, metrics
from workloads import synth_eval

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_eval_v1.npz"), allow_pickle=False)


def _sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("tag,V,seed", synth_eval.MODELS)
def test_add_adi_oracle_matches_reference(gold, tag, V, seed):
    pts = synth_eval.make_model(V, seed)
    est, gt = synth_eval.make_pose_pairs(synth_eval.N_PAIRS, seed + 1)
    assert _sha(pts, est, gt) == str(gold["err_%s_sha" % tag]), "synthetic generator drifted: regenerate the fixture"
    for i, (e, g) in enumerate(zip(est, gt)):
        a = metrics.add(e[:9].reshape(3, 3), e[9:], g[:9].reshape(3, 3), g[9:], pts)
        s = metrics.adi(e[:9].reshape(3, 3), e[9:], g[:9].reshape(3, 3), g[9:], pts)
        assert a == gold["err_%s_add" % tag][i]          # same numpy expression: bit-identical
        assert s == gold["err_%s_adi" % tag][i]
    assert gold["err_%s_add" % tag][0] == 0.0 and gold["err_%s_adi" % tag][0] == 0.0      # identical pose pair


def test_boxes_oracle_matches_reference(gold):
    boxes = synth_eval.make_boxes(synth_eval.N_BOXES, 7)
    assert _sha(boxes) == str(gold["box_sha"])
    for ratio in synth_eval.PAD_RATIOS:
        pad = np.array([evalside.padding_box(b, ratio) for b in boxes])
        assert np.array_equal(pad, gold["box_pad_%g" % ratio])
        for method in synth_eval.METHODS:
            fin = np.array([evalside.final_box(p, method, 640, 480) for p in pad])
            assert np.array_equal(fin, gold["box_final_%g_%s" % (ratio, method)]), (ratio, method)
    for method in synth_eval.METHODS:
        fin = np.array([evalside.final_box(b, method, 640, 480) for b in boxes], np.float64)
        assert np.array_equal(fin, gold["box_finalonly_%s" % method]), method
    assert evalside.final_box([1, 2, 3, 4], "something_else", 640, 480) == [1, 2, 3, 4]


def test_csv_matches_reference(gold, tmp_path):
    scene, img, Rs, ts, scores = synth_eval.make_csv_rows(9)
    want = str(gold["csv_text"])
    assert evalside.bop_csv_text(1, scene, img, Rs, ts, scores) == want
    from zebrapose_b200.tools_for_BOP import write_to_cvs
    write_to_cvs.write_cvs(str(tmp_path), "lmo_ape", 1, scene, img, Rs, ts, scores)
    assert open(tmp_path / "lmo_ape.csv").read() == want
    assert want.count("\n") == 1 + sum(1 for s in scores if s != -1)
    # batched form: same rows from [B,12] records
    poses = np.array([np.concatenate([np.asarray(r).ravel(), np.asarray(t).ravel()]) for r, t in zip(Rs, ts)])
    write_to_cvs.write_batch(str(tmp_path), "batch", 1, scene, img, poses, scores)
    assert open(tmp_path / "batch.csv").read() == want


def _crop_gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_crop_v1.npz"), allow_pickle=False)


def test_resize_restatement_matches_cv2():
    import cv2
    rng = np.random.default_rng(1)
    for t in range(120):
        sh, sw = int(rng.integers(1, 500)), int(rng.integers(1, 500))
        d = int(rng.choice([64, 128, 256]))
        if t % 10 == 0:
            sh = sw = 2 * d                      # cv2 switches to INTER_AREA for an exact 2x2 decimation
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(img, (d, d), interpolation=cv2.INTER_LINEAR), evalside.resize_linear_u8(img, d, d)), (sh, sw, d)


def test_input_crops_oracle_matches_reference():
    g = _crop_gold()
    img = synth_eval.make_image(5)
    boxes = synth_eval.make_crop_boxes(synth_eval.N_CROP_BOXES, 6)
    assert _sha(img, boxes) == str(g["in_sha"])
    for cs, method in synth_eval.CROP_CASES:
        for i, b in enumerate(boxes):
            roi = evalside.get_roi_u8(img, b, cs, method)
            assert _sha(roi) == str(g["u8_%d_%s" % (cs, method)][i]), (cs, method, i)
            assert _sha(evalside.to_tensor_normalize(roi)) == str(g["f32_%d_%s" % (cs, method)][i]), (cs, method, i)
    for i in range(3):
        roi = evalside.get_roi_u8(img, boxes[i], 64, "crop_resize")
        assert np.array_equal(roi, g["full_u8_%d" % i])
        assert np.array_equal(evalside.to_tensor_normalize(roi), g["full_f32_%d" % i])


def test_merge_csv_matches_reference(tmp_path):
    """tools_for_BOP/merge_csv.py mirror against the bytes the reference's own main() wrote for the same three result
    files (tests/golden/make_golden_merge.py): BOM, int -> float promotion of an all-integer score column, row order"""
    from zebrapose_b200.tools_for_BOP import merge_csv, write_to_cvs
    root = str(tmp_path / "results") + os.sep
    for k, (ds, obj, obj_id, seed) in enumerate(synth_eval.MERGE_FILES):
        d = os.path.join(root, ds, obj)
        os.makedirs(d)
        scene, img, Rs, ts, scores = synth_eval.make_csv_rows(seed)
        if k == 1:
            scores = [1, 1, 1, -1, 1, 1, 1]
        write_to_cvs.write_cvs(d, "%s_%s" % (ds, obj), obj_id, scene, img, Rs, ts, scores)
    out = str(tmp_path / "merged.csv")
    assert merge_csv.main(root, out) == 3
    want = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_merged_v1.csv"), "rb").read()
    assert open(out, "rb").read() == want
    assert want.startswith(b"\xef\xbb\xbfscene_id,im_id,obj_id,score,R,t,time\n") and want.count(b"\n") == 1 + 3 * 6
    with pytest.raises(ValueError):
        merge_csv.main(str(tmp_path / "nothing") + os.sep, out)
