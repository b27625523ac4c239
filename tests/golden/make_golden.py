"""Generate the golden fixtures by running the REFERENCE's own functions (imported, unmodified, from
/root/reference/zebrapose) on seeded synthetic inputs.  Runs only in the build container (the reference
mount does not exist on the GPU box); the resulting .npz / .txt files are committed.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Inputs are re-generated from seeds by oracle.synth at test time; each fixture stores a sha256 of its inputs
so a drifting generator is detected rather than silently compared against stale outputs.
"""
import hashlib
import os
import sys

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/zebrapose")

import numpy as np
import torch
from binary_code_helper.CNN_output_to_pose import (load_dict_class_id_3D_points, CNN_outputs_to_object_pose,
                                                   mapping_pixel_position_to_original_position,
                                                   build_non_unique_2D_3D_correspondence)
from binary_code_helper.class_id_encoder_decoder import class_code_images_to_class_id_image
from binary_code_helper.generate_new_dict import generate_new_corres_dict
import common_ops

from workloads import synth


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def ref_decode(mask_u8, code_f64_hwc, bbox, S, d):
    """The reference's own steps of CNN_outputs_to_object_pose up to the float32 casts (:110-129)."""
    ids = class_code_images_to_class_id_image(code_f64_hwc, 2)
    P2 = mask_u8.nonzero()
    if P2[0].size == 0:
        return ids, np.zeros((0, 2), np.float32), np.zeros((0, 3), np.float32)
    p2d, p3d = build_non_unique_2D_3D_correspondence(P2, ids, d)
    o2d = mapping_pixel_position_to_original_position(p2d, bbox, S)
    return ids, o2d.astype(np.float32), p3d.astype(np.float32)


def crop_case(tag, tab, nrm, seed, S, ignore_bit, out):
    d16 = {float(i): tab[i].copy() for i in range(len(tab))}
    c = synth.make_crop(tab, nrm, seed, S=S)
    logits = synth.crop_to_logits(c)
    lt = torch.from_numpy(logits)[None]
    # reference thresholding exactly as test.py:250-257 does it
    pm = common_ops.from_output_to_class_mask(lt[:, :1])
    pc = common_ops.from_output_to_class_binary_code(lt[:, 1:], "BCE", divided_num_each_interation=2, binary_code_length=16)
    pc = pc.transpose(0, 2, 3, 1)
    pm = pm.transpose(0, 2, 3, 1).squeeze(axis=-1).astype("uint8")
    if ignore_bit:
        d = generate_new_corres_dict(d16, 16, 16 - ignore_bit)
        code = pc[0][:, :, :-ignore_bit]
    else:
        d = d16
        code = pc[0]
    ids, uv, xyz = ref_decode(pm[0], code, c["bbox"], S, d)
    R, t, ok = CNN_outputs_to_object_pose(pm[0], code, c["bbox"], S, 2, d, intrinsic_matrix=c["K"])
    out[tag + "_in_sha"] = np.array(sha(logits, c["bbox"], tab))
    out[tag + "_ids"] = ids.astype(np.uint16)
    out[tag + "_uv"] = uv.astype(np.int16)
    assert np.array_equal(uv, uv.astype(np.int16).astype(np.float32))
    out[tag + "_xyz"] = xyz
    out[tag + "_R"] = np.asarray(R, np.float64)
    out[tag + "_t"] = np.asarray(t, np.float64)
    out[tag + "_ok"] = np.array(bool(ok))
    out[tag + "_meta"] = np.array([seed, S, ignore_bit])
    print(tag, "M=%d" % len(uv), "ok", ok)


def main():
    out = {}
    # ---- A1: dictionary text file round trip (with nan rows; one file lacking the final newline)
    tab6, _, _ = synth.make_dict(6, seed=7, radius=40.0, missing_frac=0.25)
    for nl, name in ((True, "dict_small.txt"), (False, "dict_small_nonl.txt")):
        p = os.path.join(HERE, name)
        synth.write_dict_file(p, tab6, 6, final_newline=nl)
        tot, base, nit, d = load_dict_class_id_3D_points(p)
        key = name.replace(".", "_")
        out[key + "_hdr"] = np.array([tot, base, nit])
        out[key + "_keys"] = np.array(sorted(d.keys()))
        out[key + "_vals"] = np.stack([d[k] for k in sorted(d.keys())])

    # ---- A2: ignore-bit dictionaries from the reference
    tab16n, nrm16n, _ = synth.make_dict(16, seed=11, radius=51.0, missing_frac=0.2)
    d16n = {float(i): tab16n[i].copy() for i in range(len(tab16n))}
    out["newdict_in_sha"] = np.array(sha(tab16n))
    for k in (1, 3, 8):
        nd = generate_new_corres_dict(d16n, 16, 16 - k)
        assert all(isinstance(q, int) for q in list(nd.keys())[:4]) and nd[0].shape == (1, 3)
        out["newdict_k%d" % k] = np.stack([nd[i].reshape(3) for i in range(1 << (16 - k))])

    # ---- A3: thresholds (incl. tiny magnitudes, -0.0, NaN) through the reference's torch sigmoid path
    x = np.array([-6, -1e-3, -1e-6, -0.0, 0.0, 8.9e-08, 8.9406974e-08, 1e-7, 1e-6, 1e-3, 5, np.nan, np.inf, -np.inf],
                 np.float32).reshape(1, 1, 2, 7)
    out["thr_in"] = x
    out["thr_mask"] = common_ops.from_output_to_class_mask(torch.from_numpy(x))
    out["thr_code"] = common_ops.from_output_to_class_binary_code(torch.from_numpy(x), "BCE")

    # ---- A5: pixel remap on random boxes (negative origins, non-square, S = 128 / 100 / 64)
    rng = np.random.default_rng(5)
    boxes = np.stack([rng.integers(-200, 600, 40), rng.integers(-200, 400, 40), rng.integers(1, 500, 40),
                      rng.integers(1, 500, 40)], 1).astype(np.int64)
    boxes[0] = [0, 0, 0, 0]
    boxes[1] = [-5, -5, 100, 100]
    sizes = np.array([128, 100, 64] * 14)[:40]
    res = []
    for b, S in zip(boxes, sizes):
        px = np.stack([np.arange(S), np.arange(S)[::-1]], 1)
        res.append(np.pad(mapping_pixel_position_to_original_position(px, b, int(S)), ((0, 128 - S), (0, 0))))
    out["remap_boxes"] = boxes
    out["remap_sizes"] = sizes
    out["remap_out"] = np.stack(res).astype(np.int32)
    fb = np.array([[10.5, -3.25, 77.7, 91.3], [300.0, 200.0, 64.0, 64.0]])
    px = np.stack([np.arange(128), np.arange(128)], 1)
    out["remap_fboxes"] = fb
    out["remap_fout"] = np.stack([mapping_pixel_position_to_original_position(px, b, 128) for b in fb]).astype(np.int32)

    # ---- A4/A6/A7: full crops through the reference (strict mode), S=128 and a small S=32
    tab16, nrm16, _ = synth.make_dict(16, seed=3, radius=51.0, missing_frac=0.0)
    crop_case("c1_full", tab16, nrm16, 1001 * 65536 + 0, 128, 0, out)
    crop_case("c1_nan20", tab16n, nrm16n, 1001 * 65536 + 1, 128, 0, out)
    crop_case("c3_k1", tab16n, nrm16n, 1003 * 65536 + 0, 128, 1, out)
    crop_case("c3_k4", tab16n, nrm16n, 1003 * 65536 + 1, 128, 4, out)
    crop_case("c3_k8", tab16n, nrm16n, 1003 * 65536 + 2, 128, 8, out)
    crop_case("s64_k0", tab16, nrm16, 77, 64, 0, out)
    # ignore-bit crops on the NaN-free dictionary: the only ones where the reference's POSE is meaningful
    # (with 20 % NaN rows a parent is NaN if any child is, so k=4 turns 97 % of the 3D points into (0,0,0))
    crop_case("c3f_k2", tab16, nrm16, 1003 * 65536 + 10, 128, 2, out)
    crop_case("c3f_k4", tab16, nrm16, 1003 * 65536 + 11, 128, 4, out)
    out["tab16_seed"] = np.array([3, 51.0, 0.0])
    out["tab16n_seed"] = np.array([11, 51.0, 0.2])

    # ---- edge cases of CNN_outputs_to_object_pose
    d16 = {float(i): tab16[i].copy() for i in range(len(tab16))}
    S = 128
    code = np.zeros((S, S, 16))
    m0 = np.zeros((S, S), np.uint8)
    r = CNN_outputs_to_object_pose(m0, code, np.array([0, 0, 128, 128]), S, 2, d16)
    out["edge_empty"] = np.array([len(r[0]), len(r[1]), int(r[2])])
    m5 = m0.copy(); m5[3, 3:8] = 1
    r = CNN_outputs_to_object_pose(m5, code, np.array([0, 0, 128, 128]), S, 2, d16)
    out["edge_5px"] = np.array([len(r[0]), len(r[1]), int(r[2])])
    m6 = m0.copy(); m6[3, 3:9] = 1      # 6 px, all the same code -> degenerate; reference says success=True
    r = CNN_outputs_to_object_pose(m6, code, np.array([0, 0, 128, 128]), S, 2, d16)
    out["edge_6px_ok"] = np.array(int(r[2]))
    out["edge_6px_R"] = np.asarray(r[0], np.float64)
    out["edge_6px_t"] = np.asarray(r[1], np.float64)

    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), os.path.getsize(os.path.join(HERE, "golden_v1.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
