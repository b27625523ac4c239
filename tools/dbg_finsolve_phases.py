"""Cycles of the phases of zp_fin_solve_kernel (crop 0, lane 0) via zp_debug_buffer: python tools/dbg_finsolve_phases.py [crops]"""
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
    K = torch.from_numpy(Ks.reshape(Cn, 9)).cuda()
    buf = torch.zeros(16, dtype=torch.int64, device="cuda")
    eng.ctx.check(eng.lib.zp_debug_buffer(eng.ctx.handle, C.c_void_p(buf.data_ptr())), "zp_debug_buffer")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for cold in (False, True):
        for _ in range(3):
            if cold:
                flush.zero_()
            eng.ransac(corr, counts, K)
        torch.cuda.synchronize()
        b = buf.cpu().numpy()[8:14]
        names = ["totals", "control points + A (lane 0)", "contractions", "null space (16 lanes)", "candidates (3 lanes)"]
        print("L2 %s:" % ("flushed before the call" if cold else "warm"),
              ", ".join("%s %d" % (n, b[i + 1] - b[i]) for i, n in enumerate(names)), "| total %d cycles = %.1f us" % (b[5] - b[0], (b[5] - b[0]) / 1965.0))
    eng.ctx.check(eng.lib.zp_debug_buffer(eng.ctx.handle, C.c_void_p()), "zp_debug_buffer")


if __name__ == "__main__":
    main()
