"""GPU: the reference-signature drop-ins (zebrapose_b200/binary_code_helper, common_ops) against the golden outputs
of the reference functions."""
import numpy as np
import pytest
import torch

from oracle import decode, metrics
from helpers import GOLDEN_CROPS, regen_crop

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["c1_full", "c1_nan20", "c3f_k2", "c3f_k4", "s64_k0"])
def test_cnn_outputs_to_object_pose(golden, tables, tag):
    from zebrapose_b200.binary_code_helper.CNN_output_to_pose import CNN_outputs_to_object_pose, CNN_outputs_to_object_info
    from zebrapose_b200.binary_code_helper.generate_new_dict import generate_new_corres_dict
    from zebrapose_b200 import common_ops
    tab, c, logits, S, k = regen_crop(tables, tag)
    lt = torch.from_numpy(logits)[None].cuda()
    pm = common_ops.from_output_to_class_mask(lt[:, :1])
    pc = common_ops.from_output_to_class_binary_code(lt[:, 1:], "BCE")
    pc = pc.transpose(0, 2, 3, 1)
    pm = pm.transpose(0, 2, 3, 1).squeeze(axis=-1).astype("uint8")
    d = {float(i): tab[i] for i in range(len(tab))}
    code = pc[0]
    if k:
        d = generate_new_corres_dict(d, 16, 16 - k)
        code = code[:, :, :-k]
    R, t, ok = CNN_outputs_to_object_pose(pm[0], code, c["bbox"], S, 2, d, intrinsic_matrix=torch.from_numpy(c["K"]))
    assert ok == bool(golden[tag + "_ok"]) and R.shape == (3, 3) and t.shape == (3, 1) and R.dtype == np.float64
    re, te = metrics.rot_err_deg(R, golden[tag + "_R"]), metrics.trans_err(t, golden[tag + "_t"])
    print(tag, "rot %.4f deg trans %.4f mm" % (re, te))
    # north_star tolerance on every golden crop, ignore_bit included: the hypotheses are cv2's own (exact replay)
    assert re <= 0.05 and te <= 0.5
    R2, t2, ok2, info = CNN_outputs_to_object_info(pm[0], code, c["bbox"], S, 2, d, intrinsic_matrix=c["K"])
    assert np.array_equal(R, R2) and info["n_correspondences"] == len(golden[tag + "_uv"])
    assert np.array_equal(info["coord_2d"], golden[tag + "_uv"].astype(np.float32))
    assert np.array_equal(info["coord_3d"].view(np.uint32), golden[tag + "_xyz"].view(np.uint32))


def test_edge_cases_match_reference(golden, tables):
    from zebrapose_b200.binary_code_helper.CNN_output_to_pose import CNN_outputs_to_object_pose
    tab = tables["full"][0]
    d = {float(i): tab[i] for i in range(len(tab))}
    S = 128
    code = np.zeros((S, S, 16))
    m = np.zeros((S, S), np.uint8)
    r = CNN_outputs_to_object_pose(m, code, np.array([0, 0, 128, 128]), S, 2, d)
    assert [len(r[0]), len(r[1]), int(r[2])] == golden["edge_empty"].tolist()
    m[3, 3:8] = 1
    r = CNN_outputs_to_object_pose(m, code, np.array([0, 0, 128, 128]), S, 2, d)
    assert [len(r[0]), len(r[1]), int(r[2])] == golden["edge_5px"].tolist()
    m[3, 3:9] = 1
    r = CNN_outputs_to_object_pose(m, code, np.array([0, 0, 128, 128]), S, 2, d)
    assert int(r[2]) == int(golden["edge_6px_ok"])            # success stays True
    assert np.allclose(r[0], golden["edge_6px_R"]) and np.allclose(r[1], golden["edge_6px_t"])   # R = I, t = 0


def test_small_helpers(golden):
    from zebrapose_b200.binary_code_helper.CNN_output_to_pose import mapping_pixel_position_to_original_position
    from zebrapose_b200.binary_code_helper.class_id_encoder_decoder import class_code_images_to_class_id_image
    for b, S, exp in zip(golden["remap_boxes"], golden["remap_sizes"], golden["remap_out"]):
        px = np.stack([np.arange(S), np.arange(S)[::-1]], 1)
        got = mapping_pixel_position_to_original_position(px, b, int(S))
        assert got.dtype == np.int64 and np.array_equal(got, exp[:S])
    px = np.stack([np.arange(128), np.arange(128)], 1)
    for b, exp in zip(golden["remap_fboxes"], golden["remap_fout"]):
        assert np.array_equal(mapping_pixel_position_to_original_position(px, b, 128), exp)
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2, (16, 24, 16)).astype(np.float64)
    ids = class_code_images_to_class_id_image(bits, 2)
    assert ids.dtype == np.float64 and np.array_equal(ids, decode.class_code_images_to_class_id_image(bits, 2))


def test_class_base_3_matches_reference():
    """CNN_output_to_pose.py:110 passes any class_base on; fixture = the reference's own output for a base-3, 8-digit code
    image (tests/golden/make_golden_base3.py): same correspondence lists bit for bit, pose within tolerance"""
    import os
    from zebrapose_b200.binary_code_helper.CNN_output_to_pose import CNN_outputs_to_object_info
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_base3_v1.npz"))
    d = {float(i): g["pts"][i] for i in range(len(g["pts"]))}
    R, t, ok, info = CNN_outputs_to_object_info(g["mask"], g["digits"].astype(np.float64), g["bbox"], 64, 3, d, intrinsic_matrix=g["K"])
    assert ok == bool(g["ok"])
    assert np.array_equal(info["coord_2d"], g["uv"]) and np.array_equal(info["coord_3d"].view(np.uint32), g["xyz"].view(np.uint32))
    assert metrics.rot_err_deg(R, g["R"]) <= 0.05 and metrics.trans_err(t, g["t"]) <= 0.5


def test_threshold_semantics_off_half():
    """thresholds other than 0.5 evaluate the reference's own expression (sigmoid(x) > t), incl. t <= 0 and t >= 1"""
    from zebrapose_b200 import common_ops
    x = torch.tensor([[-20.0, -1.0, -1e-8, 0.0, 1e-8, 0.5, 1.0, 20.0, float("nan")]]).reshape(1, 1, 3, 3).cuda()
    for thr in (0.0, 0.3, 0.7, 1.0, -0.5, 1.5):
        want = (torch.sigmoid(x.cpu()) > thr).to(torch.float64).numpy()
        assert np.array_equal(common_ops.from_output_to_class_mask(x, thr), want), thr
        assert np.array_equal(common_ops.from_output_to_class_binary_code(x, "BCE", thr), want), thr
