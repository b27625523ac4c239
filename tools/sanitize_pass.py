"""One small pass over every kernel family for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_pass.py      (closed on the shared pool: run natively there)
2 crops, one dictionary; progress markers on stdout so a partial log still says how far it got."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import zebrapose_b200 as zp
from workloads import synth, synth_eval


def mark(s):
    torch.cuda.synchronize()
    print("[pass]", s, flush=True)


S, B = 128, 2
tab, nrm, _ = synth.make_dict(16, seed=3, radius=51.0, missing_frac=0.1)
crops = [synth.make_crop(tab, nrm, 4242 + i, S=S) for i in range(B)]
logits = np.stack([synth.crop_to_logits(c) for c in crops])
bboxes = np.stack([c["bbox"] for c in crops]).astype(np.float64)
Ks = np.stack([c["K"] for c in crops]).reshape(B, 9)
eng = zp.Engine(0)
eng.upload_dict(0, tab)
lg = torch.from_numpy(logits).cuda()
corr, counts = eng.decode(lg, bboxes); mark("decode f32 (stream kernel) counts=%s" % counts.tolist())
eng.decode(lg.to(torch.bfloat16), bboxes); mark("decode bf16")
eng.decode(lg.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2), bboxes); mark("decode generic (channels_last strides)")
for path in (1, 3, 4, 6):
    eng.set_decode_path(path); eng.decode(lg, bboxes); mark("decode path %d" % path)
eng.set_decode_path(0)
r = eng.ransac(corr, counts, Ks); mark("ransac chain status=%s inliers=%s" % (r["status"].tolist(), r["n_inliers"].tolist()))
eng.ransac(corr, counts, Ks, sampler="philox", seed=7, select="argmax", final="epnp+gn"); mark("ransac philox / argmax / gn")
verts = np.ascontiguousarray(tab[::37][~np.isnan(tab[::37]).any(1)])
eng.upload_model(0, verts)
gt = np.stack([np.concatenate([np.asarray(c["R"], np.float64).ravel(), np.asarray(c["t"], np.float64).ravel()]) for c in crops])
eng.pose_errors(r["poses"], gt, obj_default=0); mark("ADD / ADI V=%d" % len(verts))
img = torch.from_numpy(synth_eval.make_image(5)).cuda()
cb = torch.from_numpy(synth_eval.make_crop_boxes(8, 8).astype(np.float64)).cuda()
eng.crop_inputs(img, cb, crop_size=256); eng.final_bboxes(cb, 1.5, "crop_square_resize", 640, 480); mark("crops + boxes")
g = torch.Generator(device="cpu").manual_seed(0)
eng.upload_head(torch.randn(17, 320, generator=g) * 0.1, torch.randn(17, generator=g) * 0.1)
x = torch.randn(B, 256, S, S, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
xs = torch.randn(B, 64, S, S, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
eng.head_decode(x, xs, bboxes); mark("fused head (tcgen05 + TMA) bf16")
eng.head_decode(x.float().contiguous(memory_format=torch.channels_last), xs.float().contiguous(memory_format=torch.channels_last), bboxes)
mark("fused head tf32")
print("[pass] done", flush=True)
