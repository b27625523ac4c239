"""Fused network tail vs the unfused path: python tools/bench_head.py [B ...]
unfused = torch.cat + F.conv2d (cuDNN/cuBLAS, bf16 channels_last, what the reference network does in bf16) + zp decode of
the bf16 logits;  fused = zp_head_decode.  Algorithmic HBM bytes of the fused kernel per pixel: 2*(c1+c2) in, 2.125 out."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
import zebrapose_b200 as zp
from workloads import synth

S, c1, c2 = 128, 256, 64
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
eng = zp.Engine(0)
tab, _, _ = synth.make_dict(16, seed=5, radius=60.0, missing_frac=0.0)
eng.upload_dict(0, tab)
g = torch.Generator(device="cpu").manual_seed(0)
W = (torch.randn(17, c1 + c2, generator=g) * 0.1)
bias = torch.randn(17, generator=g) * 0.1
eng.upload_head(W, bias)
Wc = W.cuda().to(torch.bfloat16).reshape(17, c1 + c2, 1, 1).contiguous(memory_format=torch.channels_last)
bc = bias.cuda().to(torch.bfloat16)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


W32c = W.cuda().reshape(17, c1 + c2, 1, 1).contiguous(memory_format=torch.channels_last)
b32c = bias.cuda()
for B in [int(v) for v in sys.argv[1:]] or [64, 256]:
    x = torch.randn(B, c1, S, S, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    xs = torch.randn(B, c2, S, S, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    bb = np.tile(np.array([[10.0, 20.0, 200.0, 200.0]]), (B, 1))
    bbt = torch.from_numpy(bb).cuda()

    def unfused():
        lg = F.conv2d(torch.cat([x, xs], 1), Wc, bc)
        return eng.decode(lg, bbt)

    def conv_only():
        return F.conv2d(torch.cat([x, xs], 1), Wc, bc)

    def fused():
        return eng.head_decode(x, xs, bbt)

    t_f, t_u, t_c = timed(fused), timed(unfused), timed(conv_only)
    corr, counts = fused()
    M = int(counts.sum())
    n_px = B * S * S
    in_bytes = n_px * 2 * (c1 + c2)
    alg = in_bytes + n_px * 2.125 + n_px * 2.125 + 20 * M + 4 * B        # head kernel in+out, emit kernel in+out
    # float32 activations: kind::tf32 tensor-core path vs torch fp32 (TF32 allowed, as a B200 deployment would set it)
    x32, xs32 = x.float().contiguous(memory_format=torch.channels_last), xs.float().contiguous(memory_format=torch.channels_last)
    torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
    t_f32 = timed(lambda: eng.head_decode(x32, xs32, bbt))
    t_u32 = timed(lambda: eng.decode(F.conv2d(torch.cat([x32, xs32], 1), W32c, b32c), bbt))
    alg32 = 2 * in_bytes + n_px * 2.125 + n_px * 2.125 + 20 * M + 4 * B
    print(json.dumps({"B": B, "dtype": "float32 activations (tf32 MMA)", "fused_us": round(t_f32, 1), "unfused_us": round(t_u32, 1),
                      "speedup": round(t_u32 / t_f32, 2), "fused_GBps": round(alg32 / t_f32 / 1e3, 1),
                      "frac_of_hbm_peak": round(alg32 / t_f32 / 1e3 / peak, 3)}))
    del x32, xs32
    print(json.dumps({"B": B, "fused_us": round(t_f, 1), "unfused_us": round(t_u, 1), "cat_conv_only_us": round(t_c, 1),
                      "speedup": round(t_u / t_f, 2), "algorithmic_bytes": int(alg), "fused_GBps": round(alg / t_f / 1e3, 1),
                      "frac_of_hbm_peak": round(alg / t_f / 1e3 / peak, 3), "hbm_peak_GBps": peak,
                      "crops_per_s_fused": round(B / t_f * 1e6)}))
