"""Cycles of the phases of stage B (zp_cvs_null_kernel) for warp 0 of CTA 0, via zp_debug_buffer."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import zebrapose_b200 as zp  # noqa: E402


def main():
    Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    eng = zp.Engine(0)
    logits, bboxes, Ks, obj, tables, crops = bench.make_workload(Cn, 1002)
    for j, t in enumerate(tables):
        eng.upload_dict(j, t, n_bits=16, ignore_bit=0)
    corr, counts = eng.decode(torch.from_numpy(logits).cuda(), bboxes, torch.from_numpy(obj.astype(np.int32)).cuda())
    s = eng.make_samples(counts, corr.shape[2], H=150, m=5)
    buf = torch.zeros(16, dtype=torch.int64, device="cuda")
    eng.ctx.check(eng.lib.zp_debug_buffer(eng.ctx.handle, C.c_void_p(buf.data_ptr())), "zp_debug_buffer")
    for H in (150, 5):
        ss = s[:, :H].contiguous()
        for _ in range(3):
            eng.solve_minimal(corr, counts, Ks, ss)
        torch.cuda.synchronize()
        b = buf.cpu().numpy()
        print("H=%d: mtm %d, jacobi %d (%d steps, %.0f cycles/step), finish %d, L/rho/out %d, total %d cycles = %.1f us" % (
            H, b[1] - b[0], b[2] - b[1], b[5], (b[2] - b[1]) / max(1, b[5]), b[3] - b[2], b[4] - b[3], b[4] - b[0], (b[4] - b[0]) / 1965.0))
    eng.ctx.check(eng.lib.zp_debug_buffer(eng.ctx.handle, C.c_void_p()), "zp_debug_buffer")


if __name__ == "__main__":
    main()
