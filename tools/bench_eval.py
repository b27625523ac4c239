"""ADD / ADI kernel timing: python tools/bench_eval.py [B V ...]  -> JSON lines (pairs/s, FP32 rate vs the measured
FFMA2 peak).  Algorithmic work of the ADI search: 9 flop per (ground-truth point, estimated point) pair
(3 sub, 1 mul + 2 fma = 5, 1 min) -> 9 V^2 per pose pair; issued as 7 packed/3-input instructions per 2 pairs."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zebrapose_b200 as zp
from workloads import synth_eval

args = [int(x) for x in sys.argv[1:]] or [64, 5841, 64, 8192, 1024, 5841, 1, 5841]
eng = zp.Engine(0)
peak = max(eng.fp32_peak_tflops(), eng.fp32_peak_tflops(packed=True))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B, V in zip(args[::2], args[1::2]):
    eng.upload_model(0, synth_eval.make_model(V, 5))
    est, gt = synth_eval.make_pose_pairs(B, 31)
    est, gt = torch.from_numpy(est).cuda(), torch.from_numpy(gt).cuda()
    res = {}
    for name, kw in (("add+adi", {}), ("add", {"adi": False})):
        for _ in range(3): eng.pose_errors(est, gt, **kw)
        tot = 0.0
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.pose_errors(est, gt, **kw); b.record(); b.synchronize()
            tot += a.elapsed_time(b)
        res[name] = tot / 10 * 1e3
    us = res["add+adi"] - res["add"]            # the search itself (prepare + finalize are in both)
    V4 = (V + 3) // 4 * 4
    pairs = float(B) * V * V4
    print(json.dumps({"B": B, "V": V, "us_add_adi": round(res["add+adi"], 2), "us_add_only": round(res["add"], 2),
                      "adi_search_us": round(us, 2), "Gpairs_per_s": round(pairs / us / 1e3, 1),
                      "tflops_9_per_pair": round(9 * pairs / us / 1e6, 2), "fp32_peak_tflops": round(peak, 1),
                      "frac_of_fp32_peak": round(9 * pairs / us / 1e6 / peak, 3),
                      "pose_pairs_per_s": round(B / res["add+adi"] * 1e6)}))
