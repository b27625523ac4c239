"""One decode + RANSAC chain on the bench workload (for ncu captures): python tools/prof_once.py [crops] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zebrapose_b200 as zp
argv = sys.argv[1:]; sys.argv = ['x']
import bench
C = int(argv[0]) if argv else 64
reps = int(argv[1]) if len(argv) > 1 else 2
logits, bboxes, Ks, obj, tables, crops = bench.make_workload(C, 1002)
eng = zp.Engine(0)
for j, t in enumerate(tables): eng.upload_dict(j, t)
lg = torch.from_numpy(logits).cuda()
for _ in range(reps):
    corr, counts = eng.decode(lg, bboxes, obj.astype(np.int32))
    r = eng.ransac(corr, counts, Ks.reshape(-1, 9))
torch.cuda.synchronize()
print("ok", int(counts.sum()))
