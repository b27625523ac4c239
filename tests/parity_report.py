#!/usr/bin/env python
"""Parity report (GPU): device decode + RANSAC-EPnP vs the reference path (oracle decode + cv2.solvePnPRansac) on the
same seeded synthetic crops.  Prints a JSON summary; run on the GPU box, copy into profiles/.
  python tests/parity_report.py [--crops 64] [--ignore-bit 0] [--out gpurun_out/parity.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cv2
import torch
import zebrapose_b200 as zp
from oracle import cvransac, decode, metrics
from workloads import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crops", type=int, default=64)
    ap.add_argument("--ignore-bit", type=int, default=0)
    ap.add_argument("--outlier", type=float, default=0.3)
    ap.add_argument("--bitflip", type=float, default=0.02)
    ap.add_argument("--seed", type=int, default=1001)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    k = args.ignore_bit
    eng = zp.Engine(0)
    tab, nrm, _ = synth.make_dict(16, seed=3, radius=51.0, missing_frac=0.0)
    eng.upload_dict(0, tab, n_bits=16, ignore_bit=k)
    tk = decode.generate_new_corres_table(tab, 16, 16 - k) if k else tab
    crops = [synth.make_crop(tab, nrm, args.seed * 65536 + i, outlier=args.outlier, bitflip=args.bitflip) for i in range(args.crops)]
    logits = np.stack([synth.crop_to_logits(c) for c in crops])
    bboxes = np.stack([c["bbox"] for c in crops])
    Ks = np.stack([c["K"] for c in crops])
    corr, counts = eng.decode(torch.from_numpy(logits).cuda(), bboxes, ignore_bit=k)
    res = eng.ransac(corr, counts, Ks, return_details=True)
    poses = res["poses"].cpu().numpy()
    best = res["best_idx"].cpu().numpy()
    iters = res["iters_run"].cpu().numpy()
    ninl = res["n_inliers"].cpu().numpy()
    hyp_inl = res["hyp_inliers"].cpu().numpy()
    im = res["inlier_mask"].cpu().numpy().astype(bool)
    corr_h, counts_h = corr.cpu().numpy(), counts.cpu().numpy()
    rot, tr, rot_gt_dev, rot_gt_ref, jac, same_w, top_rel, same_it, misses = [], [], [], [], [], [], [], [], []
    dec_ok = 0
    pts = tab[::64]
    add_dev = add_ref = 0
    for i, c in enumerate(crops):
        mask = decode.threshold_logits(logits[i, 0]).astype(np.uint8)
        code = decode.threshold_logits(logits[i, 1:]).transpose(1, 2, 0)
        if k:
            code = code[:, :, :-k]
        uv, xyz, _ = decode.decode_crop(mask, code, c["bbox"], 128, tk)
        n = counts_h[i]
        dec_ok += int(n == len(uv) and np.array_equal(corr_h[i, 0:2, :n].T, uv) and
                      np.array_equal(corr_h[i, 2:5, :n].T.view(np.uint32), xyz.view(np.uint32)))
        ok, rv, tv, inl = cv2.solvePnPRansac(xyz, uv, c["K"], None, reprojectionError=2, iterationsCount=150, flags=cv2.SOLVEPNP_EPNP)
        Rc = cv2.Rodrigues(rv)[0]
        R, t = poses[i, :9].reshape(3, 3), poses[i, 9:]
        rot.append(metrics.rot_err_deg(Rc, R)); tr.append(metrics.trans_err(tv, t))
        rot_gt_dev.append(metrics.rot_err_deg(R, c["R"])); rot_gt_ref.append(metrics.rot_err_deg(Rc, c["R"]))
        ref_mask = np.zeros(len(uv), bool)
        if inl is not None:
            ref_mask[inl.ravel()] = True
        dm = im[i, :len(uv)]
        jac.append((dm & ref_mask).sum() / max(1, (dm | ref_mask).sum()))
        _, _, _, _, info = cvransac.solve_pnp_ransac(xyz, uv, c["K"])
        same_w.append(int(best[i] == info.get("best", -2)))
        same_it.append(int(iters[i] == info["iters_run"]))
        if not (rot[-1] <= 0.05 and tr[-1] <= 0.5) or not same_w[-1]:
            misses.append({"crop": i, "rot_deg": float(rot[-1]), "trans_mm": float(tr[-1]), "winner_device": int(best[i]),
                           "winner_cv2": int(info.get("best", -2)), "inliers_device": int(ninl[i]),
                           "inliers_cv2": int(0 if inl is None else len(inl)), "n": int(len(uv))})
        cref = np.array(info["counts"]); cdev = hyp_inl[i, :len(cref)]
        good = cref >= 0.5 * cref.max()
        top_rel.append(float(np.abs(cdev[good] - cref[good]).max() / cref.max()))
        add_dev += metrics.add(R, t, c["R"], c["t"], pts) < 10.2
        add_ref += metrics.add(Rc, tv.ravel(), c["R"], c["t"], pts) < 10.2
    rot, tr = np.array(rot), np.array(tr)
    within = (rot <= 0.05) & (tr <= 0.5)
    out = {
        "crops": args.crops, "ignore_bit": k, "outlier": args.outlier, "bitflip": args.bitflip,
        "decode_bit_exact_crops": dec_ok,
        "pose_vs_cv2": {"rot_deg": {"median": float(np.median(rot)), "p90": float(np.percentile(rot, 90)), "max": float(rot.max())},
                        "trans_mm": {"median": float(np.median(tr)), "p90": float(np.percentile(tr, 90)), "max": float(tr.max())},
                        "within_0.05deg_0.5mm": int(within.sum()), "pass_rate": float(within.mean())},
        "same_winning_hypothesis": int(np.sum(same_w)),
        "same_iteration_count": int(np.sum(same_it)),
        "crops_off_tolerance_or_other_winner": misses[:20],
        "inlier_set_jaccard": {"median": float(np.median(jac)), "min": float(np.min(jac))},
        "good_hypothesis_count_rel_diff_max": {"median": float(np.median(top_rel)), "max": float(np.max(top_rel))},
        "rot_err_vs_gt_deg": {"device_median": float(np.median(rot_gt_dev)), "reference_median": float(np.median(rot_gt_ref))},
        "ADD@0.1d": {"device": int(add_dev), "reference": int(add_ref)},
    }
    print(json.dumps(out, indent=1))
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
