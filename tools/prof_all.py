"""One pass over every kernel family for ncu captures: python tools/prof_all.py [crops]
decode + RANSAC chain on the bench workload, fused network tail, ADD / ADI."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zebrapose_b200 as zp
from workloads import synth_eval
argv = sys.argv[1:]; sys.argv = ['x']
import bench
C = int(argv[0]) if argv else 64
logits, bboxes, Ks, obj, tables, crops = bench.make_workload(C, 1002)
eng = zp.Engine(0)
for j, t in enumerate(tables): eng.upload_dict(j, t)
lg = torch.from_numpy(logits).cuda()
bb = torch.from_numpy(bboxes.astype(np.float64)).cuda(); oi = torch.from_numpy(obj.astype(np.int32)).cuda()
K = torch.from_numpy(Ks.reshape(-1, 9)).cuda()
g = torch.Generator(device="cpu").manual_seed(0)
eng.upload_head(torch.randn(17, 320, generator=g) * 0.1, torch.randn(17, generator=g) * 0.1)
x = torch.randn(C, 256, 128, 128, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
xs = torch.randn(C, 64, 128, 128, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
for j in range(len(tables)): eng.upload_model(j, synth_eval.make_model(5841, 100 + j))
gt = torch.from_numpy(bench.gt_poses(crops)).cuda()
img = torch.from_numpy(synth_eval.make_image(5)).cuda()
cboxes = torch.from_numpy(synth_eval.make_crop_boxes(C, 8).astype(np.float64)).cuda()
for _ in range(2):
    corr, counts = eng.decode(lg, bb, oi)
    r = eng.ransac(corr, counts, K)
    eng.head_decode(x, xs, bb, oi)
    eng.pose_errors(r["poses"], gt, oi)
    eng.crop_inputs(img, cboxes, crop_size=256)
    eng.final_bboxes(cboxes, 1.5, "crop_square_resize", 640, 480)
torch.cuda.synchronize()
print("ok", int(counts.sum()))
