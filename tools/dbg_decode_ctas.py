"""Per-CTA timeline of the fused decode kernel (%globaltimer stamps): python tools/dbg_decode_ctas.py [crops]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import zebrapose_b200 as zp
argv = sys.argv[1:]; sys.argv = ['x']
import bench
Cn = int(argv[0]) if argv else 64
path = int(argv[1]) if len(argv) > 1 else 0
logits, bboxes, Ks, obj, tables, crops = bench.make_workload(Cn, 1002)
eng = zp.Engine(0)
for j, t in enumerate(tables): eng.upload_dict(j, t)
lg = torch.from_numpy(logits).cuda(); bb = torch.from_numpy(bboxes.astype(np.float64)).cuda(); oi = torch.from_numpy(obj.astype(np.int32)).cuda()
n_cta = Cn * 8
buf = torch.zeros(4 * n_cta, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
eng.set_decode_path(path)
for _ in range(3): eng.decode(lg, bb, oi)
eng.lib.zp_debug_buffer(eng.ctx.handle, C.c_void_p(buf.data_ptr()))
for rep in range(3):
    buf.zero_(); flush.zero_(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.decode(lg, bb, oi); b.record(); torch.cuda.synchronize()
    buf.zero_() if False else None
    t = buf.cpu().numpy().reshape(n_cta, 4).astype(np.float64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    st, mid, en = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 3] - t0) / 1e3
    print("ctas %d" % len(t), end=" ")
    print("rep %d: event %.1f us | kernel span %.1f us | CTA start p50 %.1f p90 %.1f max %.1f | lifetime p50 %.1f p90 %.1f max %.1f | load phase p50 %.1f | emit phase p50 %.1f"
          % (rep, a.elapsed_time(b) * 1e3, en.max(), np.percentile(st, 50), np.percentile(st, 90), st.max(),
             np.percentile(en - st, 50), np.percentile(en - st, 90), (en - st).max(), np.percentile(mid - st, 50), np.percentile(en - mid, 50)))
    order = np.argsort(st)
    print("   first-wave CTAs (start < 2 us): %d ; starts by decile:" % (st < 2).sum(), np.round(np.percentile(st, np.arange(0, 101, 10)), 1))
eng.lib.zp_debug_buffer(eng.ctx.handle, C.c_void_p())
