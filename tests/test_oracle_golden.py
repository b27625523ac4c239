"""CPU: the oracle restatement vs. golden outputs of the reference's own functions (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import decode, cvransac, metrics
from workloads import synth
from helpers import GOLDEN_CROPS, regen_crop, sha

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", ["dict_small.txt", "dict_small_nonl.txt"])
def test_load_dict(golden, name):
    tot, base, nit, d = decode.load_dict_class_id_3D_points(os.path.join(HERE, "golden", name))
    key = name.replace(".", "_")
    assert np.array_equal(golden[key + "_hdr"], [tot, base, nit])
    ks = sorted(d.keys())
    assert np.array_equal(golden[key + "_keys"], ks)
    assert all(isinstance(k, float) for k in ks)
    np.testing.assert_array_equal(golden[key + "_vals"], np.stack([d[k] for k in ks]))  # NaN == NaN here


@pytest.mark.parametrize("k", [1, 3, 8])
def test_generate_new_dict(golden, tables, k):
    tab = tables["nan20"][0]
    assert sha(tab) == str(golden["newdict_in_sha"])
    new = decode.generate_new_corres_table(tab, 16, 16 - k)
    np.testing.assert_array_equal(golden["newdict_k%d" % k], new)      # bit-exact incl. NaN rows
    nd = decode.generate_new_corres_dict({float(i): tab[i] for i in range(len(tab))}, 16, 16 - k)
    assert isinstance(next(iter(nd.keys())), int) and nd[0].shape == (1, 3)


def test_threshold(golden):
    x = golden["thr_in"]
    # product rule float32(x) > 0 differs from sigmoid(x) > 0.5 only for 0 < x < 8.9406974e-08
    ours = decode.threshold_logits(x)
    ref = golden["thr_mask"]
    tiny = (x > 0) & (x < np.float32(8.9406974e-08))
    assert np.array_equal(ours[~tiny], ref[~tiny])
    assert np.array_equal(golden["thr_code"], ref)


def test_pixel_remap(golden):
    for b, S, exp in zip(golden["remap_boxes"], golden["remap_sizes"], golden["remap_out"]):
        px = np.stack([np.arange(S), np.arange(S)[::-1]], 1)
        got = decode.mapping_pixel_position_to_original_position(px, b, int(S))
        assert np.array_equal(got, exp[:S])
    px = np.stack([np.arange(128), np.arange(128)], 1)
    for b, exp in zip(golden["remap_fboxes"], golden["remap_fout"]):
        assert np.array_equal(decode.mapping_pixel_position_to_original_position(px, b, 128), exp)


@pytest.mark.parametrize("tag", list(GOLDEN_CROPS))
def test_decode_crop(golden, tables, tag):
    tab, c, logits, S, k = regen_crop(tables, tag)
    assert sha(logits, c["bbox"], tab) == str(golden[tag + "_in_sha"]), "synthetic generator drifted"
    mask = decode.threshold_logits(logits[0]).astype(np.uint8)
    code = decode.threshold_logits(logits[1:]).transpose(1, 2, 0)
    t = tab
    if k:
        t = decode.generate_new_corres_table(tab, 16, 16 - k)
        code = code[:, :, :-k]
    uv, xyz, ids = decode.decode_crop(mask, code, c["bbox"], S, t)
    assert np.array_equal(ids.astype(np.uint16), golden[tag + "_ids"])
    assert np.array_equal(uv, golden[tag + "_uv"].astype(np.float32))
    assert np.array_equal(xyz.view(np.uint32), golden[tag + "_xyz"].view(np.uint32))
    # faithful per-pixel loop == vectorised
    d = decode.table_to_dict(t, float_keys=True)
    p2, p3 = decode.build_correspondences_faithful(mask, decode.class_code_images_to_class_id_image(code), d)
    assert np.array_equal(p3.astype(np.float32), xyz)


@pytest.mark.parametrize("tag", ["c1_full", "c3f_k4", "s64_k0"])
def test_pose_vs_reference(golden, tables, tag):
    """oracle decode + RANSAC emulation (cv2 EPnP inside) == the reference's CNN_outputs_to_object_pose."""
    tab, c, logits, S, k = regen_crop(tables, tag)
    mask = decode.threshold_logits(logits[0]).astype(np.uint8)
    code = decode.threshold_logits(logits[1:]).transpose(1, 2, 0)
    t = decode.generate_new_corres_table(tab, 16, 16 - k) if k else tab
    if k:
        code = code[:, :, :-k]
    uv, xyz, _ = decode.decode_crop(mask, code, c["bbox"], S, t)
    ok, R, tv, inl, info = cvransac.solve_pnp_ransac(xyz, uv, c["K"])
    assert ok == bool(golden[tag + "_ok"])
    assert metrics.rot_err_deg(R, golden[tag + "_R"]) < 1e-4
    assert metrics.trans_err(tv, golden[tag + "_t"]) < 1e-3


def test_hamming_remap_spec():
    rng = np.random.default_rng(0)
    exists = rng.random(256) > 0.4
    remap = decode.hamming_remap_table(exists)
    for c in range(256):
        best = min((bin(c ^ e).count("1"), c ^ e, e) for e in np.nonzero(exists)[0])
        assert remap[c] == best[2]
    assert np.array_equal(remap[exists], np.nonzero(exists)[0])
