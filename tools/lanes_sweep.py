"""Device-resident throughput (poses/s) of decode + RANSAC against the number of lanes (batches in flight), per solver.
python tools/lanes_sweep.py [crops] [steps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import zebrapose_b200 as zp  # noqa: E402


def main():
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    logits, bboxes, Ks, obj, tables, crops = bench.make_workload(C, 1002)
    n_buf = 6 if C <= 128 else 2
    for solver in ("cv2", "fast"):
        for lanes in (1, 2, 3, 4, 6, 8):
            pipe = zp.Pipeline(0, lanes=lanes)
            for e in pipe.engines:
                e.set_solver(solver)
            for j, t in enumerate(tables):
                pipe.upload_dict(j, t, n_bits=16, ignore_bit=0)
            bufs = []
            for j in range(n_buf):
                r = (j * 7) % C
                bufs.append((torch.from_numpy(np.roll(logits, r, 0)).cuda(), torch.from_numpy(np.roll(bboxes, r, 0).astype(np.float64)).cuda(),
                             torch.from_numpy(np.roll(Ks.reshape(C, 9), r, 0)).cuda(), torch.from_numpy(np.roll(obj, r, 0).astype(np.int32)).cuda()))
            kw = dict(n_bits=16, iters=150, m=5, thr=2.0)
            for i in range(max(6, lanes)):
                pipe.submit(*bufs[i % n_buf], **kw)
            pipe.join(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                pipe.submit(*bufs[i % n_buf], **kw)
            pipe.join()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print(json.dumps({"crops": C, "solver": solver, "lanes": lanes, "ms_per_step": round(ms / steps, 4),
                              "poses_per_s": round(C * steps / ms * 1e3)}), flush=True)
            del pipe
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
