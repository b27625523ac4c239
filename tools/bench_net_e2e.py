"""BASELINE.json configs[4] on the GPUs of one box: random-init ZebraPose network, bf16 forward on 256 x 256 crops, feeding
the device-side pose path without a host copy.

    python tools/bench_net_e2e.py [crops_per_gpu=128] [steps=10]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_net_e2e.py 128

Arms (CUDA events on the current stream, 3 warm-up steps, the step's own GBs of activations flush L2 between steps):
  body      the network up to its last activations (torch / cuDNN, channels_last, BatchNorms folded) -- NOT this
            repo's code, timed so the path's share of the step is visible
  fused     body -> zp_head_decode (conv_1x1_4 on tcgen05 + threshold + pack + emit) -> RANSAC chain      [the product]
  unfused   body -> torch conv_1x1_4 -> 17-plane logits -> zp_decode -> RANSAC chain      [what the reference graph does]
Random-init logits are noise: ~95 % of the pixels are "masked" and no hypothesis gathers inliers, so
RANSAC runs all 150 iterations on ~15.6 k correspondences per crop -- its worst case; poses are meaningless.
N ranks: weak scaling, crops_per_gpu each, one final all_gather of the poses inside the timed region, max over ranks."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from workloads import net as znet
from workloads import synth


def measure(eng, B, steps=10, warmup=3, dist=None):
    dev = eng.device
    net = znet.build(seed=0, device=dev, dtype=torch.bfloat16, fold=True)
    img = znet.images(B, seed=0, device=dev, dtype=torch.bfloat16)
    tail_w = net.tail.weight.detach().float().reshape(17, 320)
    tail_b = net.tail.bias.detach().float()
    eng.upload_head(tail_w, tail_b)
    bb = torch.tensor([[100.0, 60.0, 180.0, 180.0]], device=dev, dtype=torch.float64).repeat(B, 1)
    K = torch.from_numpy(np.tile(synth.YCBV_K.reshape(1, 9), (B, 1))).to(dev)
    obj = torch.zeros(B, dtype=torch.int32, device=dev)
    world = dist.get_world_size() if dist is not None else 1

    def body():
        with torch.no_grad():
            return net(img)

    def fused():
        x, xs = body()
        return eng.head_pose_batch(x, xs, bb, K, obj)

    def unfused():
        x, xs = body()
        with torch.no_grad():
            lg = net.tail(torch.cat([x, xs], 1)).contiguous()            # NCHW planes, as the reference network returns them
        return eng.decode_and_pose_batch(lg, bb, K, obj)

    def pose_only(x, xs):
        return eng.head_pose_batch(x, xs, bb, K, obj)

    def timed(fn, gather=False):
        for _ in range(warmup):
            out = fn()
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            out = fn()
        if gather and dist is not None:
            rec = torch.cat([out[0], out[1].double().unsqueeze(1), out[2].double().unsqueeze(1)], 1)
            allr = torch.empty((world * B, 14), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allr, rec)
        b.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    t_body, (x, xs) = timed(body)
    in_place = bool(x.permute(0, 2, 3, 1).is_contiguous() and xs.permute(0, 2, 3, 1).is_contiguous())
    t_pose, _ = timed(lambda: pose_only(x, xs))
    n0 = eng.launch_count()
    t_fused, out = timed(fused, gather=True)
    launches = (eng.launch_count() - n0) // (steps + warmup)
    t_unf, _ = timed(unfused, gather=True)
    _, cnt = eng.head_decode(x, xs, bb, obj)
    flop = 109.1e9 * B                                                   # reference network, 2 * MAC
    return {"workload": "configs[4]: random-init ResNet34-OS8 + ASPP body (torch/cuDNN, bf16 channels_last, folded BatchNorm) on "
                        "%d x 3 x 256 x 256 N(0,1) crops per GPU -> fused head -> decode -> RANSAC-EPnP (150 it, 2 px)" % B,
            "n_gpus": world, "crops_per_gpu": B, "steps": steps, "warmup": warmup,
            "ms_body": round(t_body, 3), "ms_pose_path_alone": round(t_pose, 3),
            "ms_step_fused": round(t_fused, 3), "ms_step_unfused": round(t_unf, 3),
            "crops_per_s_fused": round(world * B / t_fused * 1e3, 1), "crops_per_s_unfused": round(world * B / t_unf * 1e3, 1),
            "pose_path_share_of_step": round(1.0 - t_body / t_fused, 4),
            "body_tflops_bf16": round(flop / t_body / 1e9, 1),
            "activations_consumed_in_place": in_place, "zp_launches_per_step": int(launches),
            "masked_px_per_crop": round(float(cnt.float().mean().item()), 1),
            "status_counts": {int(k): int(v) for k, v in zip(*np.unique(out[2].cpu().numpy(), return_counts=True))},
            "timing": "CUDA events around %d steps after %d warm-up, max over ranks; final pose all_gather inside" % (steps, warmup)}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    import zebrapose_b200 as zp
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)                       # NCCL prints its version banner on stdout (at init with device_id, or at the first collective)
    try:
        if world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        eng = zp.Engine(local)
        tab, _, _ = synth.make_dict(16, seed=5, radius=60.0, missing_frac=0.0)
        eng.upload_dict(0, tab)
        row = measure(eng, B, steps=steps, dist=dist)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(row))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
