"""CPU enqueue cost of one pipelined step vs GPU time: python tools/cpu_overhead.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zebrapose_b200 as zp
sys.argv = ['x']
import bench
C = 64
logits, bboxes, Ks, obj, tables, crops = bench.make_workload(C, 1002)
pipe = zp.Pipeline(0, lanes=3)
for j, t in enumerate(tables): pipe.upload_dict(j, t)
lg = torch.from_numpy(logits).cuda(); bb = torch.from_numpy(bboxes.astype(np.float64)).cuda()
K = torch.from_numpy(Ks.reshape(C, 9)).cuda(); oi = torch.from_numpy(obj.astype(np.int32)).cuda()
for _ in range(6): pipe.submit(lg, bb, K, oi)
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for _ in range(n): pipe.submit(lg, bb, K, oi)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("cpu enqueue %.1f us/step; total %.1f us/step" % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
