"""decode kernel timing per path: python tools/bench_decode.py [crops]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zebrapose_b200 as zp
argv = sys.argv[1:]; sys.argv = ['x']
import bench
for C in [int(x) for x in argv] or [64, 1024]:
    logits, bboxes, Ks, obj, tables, crops = bench.make_workload(C, 1002)
    eng = zp.Engine(0)
    for j, t in enumerate(tables): eng.upload_dict(j, t)
    for dt in (torch.float32, torch.bfloat16):
        lg = torch.from_numpy(logits).cuda().to(dt)
        bb = torch.from_numpy(bboxes.astype(np.float64)).cuda(); oi = torch.from_numpy(obj.astype(np.int32)).cuda()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        flush2 = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
        clean = len(os.environ.get("ZP_CLEAN_FLUSH", "")) > 0
        for path in (0,) if os.environ.get('ZP_DECODE_ONLY_DEFAULT') else (0, 6, 3, 4, 102, 104):
            eng.set_decode_path(path)
            for _ in range(3): corr, counts = eng.decode(lg, bb, oi)
            tot = 0.0
            for _ in range(20):
                flush.zero_()
                if clean: flush2.sum()      # leaves L2 full of CLEAN lines: no dirty write-back inside the timed kernel
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); corr, counts = eng.decode(lg, bb, oi); b.record(); b.synchronize()
                tot += a.elapsed_time(b)
            us = tot / 20 * 1e3
            M = int(counts.sum())
            byts = C * 17 * 128 * 128 * lg.element_size() + 20 * M + 4 * C
            print("crops %5d %s path %d: %8.2f us  %7.1f GB/s  (%.1f%% of 6546.6)" % (C, str(dt)[6:], path, us, byts / us / 1e3, byts / us / 1e3 / 65.466))
        eng.set_decode_path(100); eng.set_decode_path(0)
