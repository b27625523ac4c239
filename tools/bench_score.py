"""score kernel timing per groups-per-CTA setting: python tools/bench_score.py [crops ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zebrapose_b200 as zp
argv = sys.argv[1:]; sys.argv = ['x']
import bench
print('ZP_SCORE_PER_SM', os.environ.get('ZP_SCORE_PER_SM'))
for C in [int(x) for x in argv] or [64, 1024]:
    logits, bboxes, Ks, obj, tables, crops = bench.make_workload(C, 1002)
    eng = zp.Engine(0)
    for j, t in enumerate(tables): eng.upload_dict(j, t)
    lg = torch.from_numpy(logits).cuda()
    corr, counts = eng.decode(lg, bboxes, obj.astype(np.int32))
    K = torch.from_numpy(Ks.reshape(C, 9)).cuda()
    cap = corr.shape[2]
    samples = eng.make_samples(counts, cap, 150, 5)
    hyp = eng.solve_minimal(corr, counts, K, samples)
    M = int(counts.clamp(max=cap).sum())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peak = max(eng.fp32_peak_tflops(), eng.fp32_peak_tflops(packed=True))
    for g, hc in ((1, -1), (1, 100), (1, 90), (1, 75), (1, 60), (1, 50), (1, 30), (0, 0)):
        eng.set_score_groups(g, hc)
        for _ in range(3): eng.score(corr, counts, K, hyp, 2.0)
        tot = 0.0
        for _ in range(20):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.score(corr, counts, K, hyp, 2.0); b.record(); b.synchronize()
            tot += a.elapsed_time(b)
        us = tot / 20 * 1e3
        tf = 27.0 * 150 * M / us / 1e6
        print("crops %5d groups %d chunk %3d: %8.2f us  %5.1f TFLOP/s  (%.1f%% of %.1f)" % (C, g, hc, us, tf, 100 * tf / peak, peak))
