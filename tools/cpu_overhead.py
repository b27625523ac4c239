"""Host enqueue cost of one pipelined step (CPU time spent in Pipeline.submit) against the GPU time per step, for the
eager chain (10+ launches through ctypes) and the CUDA-graph replay (one cudaGraphLaunch).
python tools/cpu_overhead.py [crops] [lanes] > profiles/r2_cpu_overhead.jsonl"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zebrapose_b200 as zp

argv, sys.argv = sys.argv, ["x"]
import bench  # noqa: E402

C = int(argv[1]) if len(argv) > 1 else 64
lanes = int(argv[2]) if len(argv) > 2 else 6
logits, bboxes, Ks, obj, tables, crops = bench.make_workload(C, 1002)
pipe = zp.Pipeline(0, lanes=lanes)
for j, t in enumerate(tables):
    pipe.upload_dict(j, t)
bufs = []
for j in range(lanes):                       # one fixed input set per lane: the graph of a lane names its own buffers
    r = (j * 7) % C
    bufs.append((torch.from_numpy(np.roll(logits, r, 0)).cuda(), torch.from_numpy(np.roll(bboxes, r, 0).astype(np.float64)).cuda(),
                 torch.from_numpy(np.roll(Ks.reshape(C, 9), r, 0)).cuda(), torch.from_numpy(np.roll(obj, r, 0).astype(np.int32)).cuda()))
for graph in (False, True):
    for i in range(4 * lanes):
        pipe.submit(*bufs[i % lanes], graph=graph)
    pipe.join(); torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter()
    for i in range(n):
        pipe.submit(*bufs[i % lanes], graph=graph)
    t1 = time.perf_counter()
    pipe.join(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(json.dumps({"crops": C, "lanes": lanes, "graph": graph, "host_enqueue_us_per_step": round((t1 - t0) / n * 1e6, 1),
                      "wall_us_per_step": round((t2 - t0) / n * 1e6, 1), "poses_per_s": round(C * n / (t2 - t0))}), flush=True)
