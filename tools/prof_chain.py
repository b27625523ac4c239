"""decode + RANSAC chain only, twice, for ncu captures: python tools/prof_chain.py [crops] [one-wave: 0|1]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zebrapose_b200 as zp
argv = sys.argv[1:]; sys.argv = ['x']
import bench
C = int(argv[0]) if argv else 64
logits, bboxes, Ks, obj, tables, crops = bench.make_workload(min(C, 64), 1002)
rep = max(1, C // 64)
eng = zp.Engine(0)
for j, t in enumerate(tables): eng.upload_dict(j, t)
if len(argv) > 1 and int(argv[1]): eng.set_waves([150])
lg = torch.from_numpy(logits).cuda().repeat(rep, 1, 1, 1)
bb = torch.from_numpy(bboxes.astype(np.float64)).cuda().repeat(rep, 1); oi = torch.from_numpy(obj.astype(np.int32)).cuda().repeat(rep)
K = torch.from_numpy(Ks.reshape(-1, 9)).cuda().repeat(rep, 1)
for _ in range(2):
    corr, counts = eng.decode(lg, bb, oi)
    r = eng.ransac(corr, counts, K)
torch.cuda.synchronize()
print("ok", int(counts.sum()), int(r["n_inliers"].sum()))
