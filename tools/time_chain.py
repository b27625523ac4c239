"""Per-kernel durations of the RANSAC chain (library event pairs, L2 flushed) for both minimal solvers and several wave
plans; plus the iteration counts cv2's adaptive stop leaves.  python tools/time_chain.py [crops ...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import zebrapose_b200 as zp  # noqa: E402

NAMES = ["zp_samples_kernel", "zp_cvs_prep_kernel", "zp_cvs_null_kernel", "zp_cvs_cand_kernel", "zp_cvs_pick_kernel", "zp_minimal_kernel", "zp_score_kernel", "zp_rs_replay_kernel",
         "zp_fin_moments_kernel", "zp_fin_solve_kernel", "zp_fin_errors_kernel", "zp_final_kernel"]


def main():
    crops_list = [int(a) for a in sys.argv[1:]] or [64, 1024]
    eng = zp.Engine(0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for C in crops_list:
        if os.environ.get("TC_CLEAN"):          # 30 % outliers, no bit flips: cv2 stops after ~30 iterations
            from workloads import synth
            logits, bboxes, Ks, obj, tables, crops = synth.make_batch(min(C, 64), S=128, n_bits=16, n_dicts=21, seed=1002, K=synth.YCBV_K,
                                                                      outlier=0.3, bitflip=0.0, radius=(40.0, 175.0))
            rep = C // len(logits)
            logits, bboxes, Ks, obj = np.tile(logits, (rep, 1, 1, 1)), np.tile(bboxes, (rep, 1)), np.tile(Ks, (rep, 1, 1)), np.tile(obj, rep)
        else:
            logits, bboxes, Ks, obj, tables, crops = bench.make_workload(C, 1002)
        for j, t in enumerate(tables):
            eng.upload_dict(j, t, n_bits=16, ignore_bit=0)
        d_logits = torch.from_numpy(logits).cuda()
        d_obj = torch.from_numpy(obj.astype(np.int32)).cuda()
        corr, counts = eng.decode(d_logits, bboxes, d_obj)
        plans = (("cv2", [150]), ("cv2", None), ("cv2", [32]), ("cv2", [32, 118]), ("fast", [150]), ("fast", None))
        if os.environ.get("TC_QUICK"):
            plans = plans[:1]
        if os.environ.get("TC_WAVES"):
            plans = (("cv2", [150]), ("cv2", None), ("cv2", [32, 118]), ("cv2", [32]))
        for solver, plan in plans:
            eng.set_solver(solver)
            eng.set_waves(plan)
            fn = lambda: eng.ransac(corr, counts, Ks, H=150, m=5, thr=2.0)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            tot = []
            for _ in range(10):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); b.synchronize()
                tot.append(a.elapsed_time(b))
            eng.set_kernel_timing(True)
            for _ in range(10):
                flush.zero_()
                fn()
            torch.cuda.synchronize()
            k = {n: eng.kernel_time(n) for n in NAMES}
            eng.set_kernel_timing(False)
            row = {"env": {k: v for k, v in os.environ.items() if k.startswith("ZP_")}, "crops": C, "solver": solver, "waves": plan, "chain_ms_median": round(float(np.median(tot)), 4),
                   "kernel_us_per_call": {n: round(v[0] * v[1] * 1e3 / 10, 1) for n, v in k.items() if v[1]},
                   "launches_per_call": {n: v[1] // 10 for n, v in k.items() if v[1]}}
            print(json.dumps(row), flush=True)
        eng.set_solver("cv2"); eng.set_waves(None)
        r = eng.ransac(corr, counts, Ks, return_details="state")
        it = r["iters_run"].cpu().numpy()
        print(json.dumps({"crops": C, "iters_run": {"min": int(it.min()), "median": float(np.median(it)), "p90": float(np.percentile(it, 90)),
                                                     "max": int(it.max()), "mean": float(it.mean())}}), flush=True)


if __name__ == "__main__":
    main()
