#!/usr/bin/env python
"""bench.py -- crop-poses/s of the post-network pose path (decode + RANSAC-PnP) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--crops C]

A step = one pass of the hot path over one batch of C synthetic crops per GPU (default: BASELINE.json configs[1],
64 YCB-V-like 128x128 crops, 21 dictionaries, ignore_bit 0).  `value` = whole-job poses/s with the logits already
resident in HBM, K steps enqueued round-robin on --lanes contexts/streams and timed as one region (`step_latency_ms` is
one step alone); `e2e` = the same through zp_pose_batch_host_async/zp_sync (HOST pinned buffers, H2D + D2H inside the
timed region).
`--impl reference` times the reference's own CPU path (oracle/reference_path.py: restated per-pixel dict loop +
cv2.solvePnPRansac) on all host cores.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "crop-poses/sec, code decode+RANSAC-PnP"
UNIT = "poses/s"
S, NBITS, H, M, THR = 128, 16, 150, 5, 2.0
L2_FLUSH_BYTES = 256 << 20
WORKLOAD = ("configs[1]: %d synthetic YCB-V-like 128x128 crops per GPU, 21 dictionaries, ignore_bit 0, "
            "decode + RANSAC-EPnP (150 hypotheses x 5 points, 2 px, cv2 replay)")


def make_workload(crops, seed):
    from workloads import synth
    return synth.make_batch(crops, S=S, n_bits=NBITS, n_dicts=21, seed=seed, K=synth.YCBV_K, outlier=0.3, bitflip=0.02,
                            missing_frac=0.0, radius=(40.0, 175.0))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


class StdoutToStderr:
    """fd-level redirect: libraries that write to stdout (NCCL's version banner) must not pollute the one JSON line"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def bind_to_gpu_numa(gpu):
    """Pin this rank to the CPUs NVML reports as local to its GPU BEFORE the pinned host buffers are allocated and first
    touched, so that the pages of the e2e path sit on the GPU's NUMA node (a remote-socket buffer costs ~25 % of the
    host->device rate).  Returns the previous affinity (restored before the CPU baseline runs on all cores)."""
    try:
        prev = os.sched_getaffinity(0)
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & prev
        if cpus:
            os.sched_setaffinity(0, cpus)
        return prev, sorted(cpus)
    except Exception:
        return None, None


def cpu_baseline(logits, bboxes, Ks, obj, tables, n_sample, repeats=1):
    """reference CPU path on all host cores, bounded sample"""
    from oracle.reference_path import ReferencePool
    cores = os.cpu_count() or 1
    pool = ReferencePool(logits, bboxes, Ks, obj, tables, cores)
    idx = [i % len(logits) for i in range(n_sample)]
    pool.run(idx[: 2 * cores])                      # warm-up map
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        pool.run(idx)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    poses = pool.run(list(range(len(logits))))      # the batch itself once more (untimed): poses for the agreement figures
    pool.close()
    return n_sample / best, cores, best, poses


def run_reference(args, rank, world):
    if rank != 0:
        return
    logits, bboxes, Ks, obj, tables, _ = make_workload(args.crops, 1002)
    from oracle.reference_path import ReferencePool
    cores = os.cpu_count() or 1
    pool = ReferencePool(logits, bboxes, Ks, obj, tables, cores)
    idx = list(range(args.crops))
    for _ in range(args.warmup):
        pool.run(idx[: 2 * cores])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.run(idx)
    dt = time.perf_counter() - t0
    pool.close()
    val = args.crops * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % args.crops, "crops_per_gpu": args.crops,
                       "arm": "reference CPU path (host threshold of all logits, per-pixel dict loop, cv2.solvePnPRansac EPnP), "
                              "one step = the same batch on all host cores"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d crops per step, fork pool of %d workers, cv2.setNumThreads(1)" % (args.crops, cores)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def gt_poses(crops):
    import numpy as np
    return np.stack([np.concatenate([np.asarray(c["R"], np.float64).ravel(), np.asarray(c["t"], np.float64).ravel()]) for c in crops])


def agreement(eng, d_logits, d_bbox, d_K, d_obj, crops, tables, ref_poses, kw):
    """This batch through the device path vs the reference CPU path (cv2.solvePnPRansac), crop by crop: pose differences
    and ADD@0.1d against ground truth for both (device ADD kernel; model = every 11th dictionary point; d = its extent)."""
    import numpy as np
    import torch
    from oracle import metrics
    ours = eng.decode_and_pose_batch(d_logits, d_bbox, d_K, d_obj, **kw)[0].cpu().numpy()
    gt = gt_poses(crops)
    obj = d_obj.cpu().numpy()
    diam = np.zeros(len(tables))
    for j, t in enumerate(tables):
        v = np.ascontiguousarray(t[::11])
        eng.upload_model(j, v)
        sub = v[:: max(1, len(v) // 800)]
        diam[j] = np.sqrt(((sub[:, None, :] - sub[None, :, :]) ** 2).sum(-1).max())
    g = torch.from_numpy(gt).cuda()
    add_o = eng.pose_errors(torch.from_numpy(ours).cuda(), g, d_obj, adi=False)[0].cpu().numpy()
    add_r = eng.pose_errors(torch.from_numpy(np.ascontiguousarray(ref_poses)).cuda(), g, d_obj, adi=False)[0].cpu().numpy()
    thr = 0.1 * diam[obj]
    rot = np.array([metrics.rot_err_deg(a[:9].reshape(3, 3), b[:9].reshape(3, 3)) for a, b in zip(ours, ref_poses)])
    tr = np.array([metrics.trans_err(a[9:], b[9:]) for a, b in zip(ours, ref_poses)])
    return {"crops": int(len(ours)), "add_0.1d_pass_ours": float((add_o < thr).mean()), "add_0.1d_pass_reference": float((add_r < thr).mean()),
            "add_0.1d_same_verdict": float(((add_o < thr) == (add_r < thr)).mean()),
            "pose_within_0.05deg_0.5mm": float(((rot <= 0.05) & (tr <= 0.5)).mean()),
            "rot_diff_deg_median": float(np.median(rot)), "rot_diff_deg_p90": float(np.percentile(rot, 90)),
            "trans_diff_mm_median": float(np.median(tr)), "trans_diff_mm_p90": float(np.percentile(tr, 90)),
            "add_mm_median_ours": float(np.median(add_o)), "add_mm_median_reference": float(np.median(add_r))}


def next_rows(eng, timed, d_bbox, d_K, d_obj, crops, tables, C, hbm_peak, fp32_peak):
    """SURVEY 8(f) rows measured beside the headline (same crops, rank 0, CUDA events, L2 flushed): the fused network tail
    (N1: 1x1 conv on tcgen05 + threshold + pack + emit, random bf16 activations of the reference's 256+64 channels) and
    ADD / ADI of the batch's poses against ground truth (N3; 5841-vertex models = LM-O ape size)."""
    import numpy as np
    import torch
    from workloads import synth_eval
    out = {}
    g = torch.Generator(device="cpu").manual_seed(0)
    c1, c2 = 256, 64
    eng.upload_head(torch.randn(17, c1 + c2, generator=g) * 0.1, torch.randn(17, generator=g) * 0.1)
    x = torch.randn(C, c1, S, S, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    xs = torch.randn(C, c2, S, S, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ms = timed(lambda: eng.head_decode(x, xs, d_bbox, d_obj), reps=10)
    _, cnt = eng.head_decode(x, xs, d_bbox, d_obj)
    n_px = C * S * S
    alg = n_px * 2 * (c1 + c2) + 2 * n_px * 2.125 + 20 * int(cnt.sum().item()) + 4 * C
    out["head_decode"] = {"us": round(ms * 1e3, 1), "algorithmic_bytes": int(alg), "GBps": round(alg / ms / 1e6, 1),
                          "frac_of_hbm_peak": round(alg / ms / 1e6 / hbm_peak, 3), "crops_per_s": round(C / ms * 1e3),
                          "kernels": "zp_head_codes_kernel (tcgen05 + TMA) + zp_decode_emit_kernel"}
    del x, xs
    # N2: the network's input crops from a device-resident 480 x 640 frame (get_roi + ToTensor + Normalize), fp32 NCHW as
    # the reference feeds its network and bf16 channels_last as a B200 network wants them
    img = torch.from_numpy(synth_eval.make_image(5)).cuda()
    boxes = torch.from_numpy(synth_eval.make_crop_boxes(C, 8).astype(np.float64)).cuda()
    ms32 = timed(lambda: eng.crop_inputs(img, boxes, crop_size=256), reps=10)
    ms16 = timed(lambda: eng.crop_inputs(img, boxes, crop_size=256, dtype=torch.bfloat16, channels_last=True), reps=10)
    out["input_crops"] = {"us_f32_nchw": round(ms32 * 1e3, 1), "us_bf16_channels_last": round(ms16 * 1e3, 1),
                          "GBps_written_f32": round(C * 3 * 256 * 256 * 4 / ms32 / 1e6, 1),
                          "frac_of_hbm_peak_f32": round(C * 3 * 256 * 256 * 4 / ms32 / 1e6 / hbm_peak, 3),
                          "crops_per_s_f32": round(C / ms32 * 1e3)}
    V = 5841
    n_obj = len(tables)
    for j in range(n_obj):
        eng.upload_model(j, synth_eval.make_model(V, 100 + j))
    gt = torch.from_numpy(gt_poses(crops)).cuda()
    est = gt.clone(); est[:, 9:] += 0.5                      # poses 0.5 mm off: the search sees realistic near-coincident sets
    ms_e = timed(lambda: eng.pose_errors(est, gt, d_obj), reps=10)
    ms_a = timed(lambda: eng.pose_errors(est, gt, d_obj, adi=False), reps=10)
    pairs = float(C) * V * ((V + 3) // 4 * 4)
    us = (ms_e - ms_a) * 1e3
    out["pose_errors"] = {"us_add_adi": round(ms_e * 1e3, 1), "us_add_only": round(ms_a * 1e3, 1), "vertices": V,
                          "adi_tflops_9_per_pair": round(9 * pairs / us / 1e6, 2),
                          "adi_frac_of_fp32_peak": round(9 * pairs / us / 1e6 / fp32_peak, 3) if fp32_peak else None,
                          "pose_pairs_per_s": round(C / ms_e * 1e3)}
    # configs[4]: the random-init network's bf16 forward feeding the path on the device (body = torch / cuDNN, not this
    # repo's code; reported so the path's share of an end-to-end step is on record).  128 crops = 1024 over 8 GPUs.
    try:
        from tools import bench_net_e2e
        out["network_feed"] = bench_net_e2e.measure(eng, 128, steps=5, warmup=3)
    except Exception as exc:
        out["network_feed"] = {"error": repr(exc)[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--crops", type=int, default=64, help="crops per GPU per step (configs[1] = 64)")
    ap.add_argument("--lanes", type=int, default=3, help="batches in flight per GPU (one zp_ctx + CUDA stream each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernels", action="store_true", help="also print a per-kernel table to stderr")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import zebrapose_b200 as zp

    args.warmup = max(args.warmup, 3)
    prev_affinity, numa_cpus = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with StdoutToStderr():
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
    pipe = zp.Pipeline(local, lanes=args.lanes)
    eng = pipe.engines[0]
    C = args.crops
    # weak scaling: every rank owns its own batch of C crops (contiguous shard [rank*C, (rank+1)*C) of the job)
    logits, bboxes, Ks, obj, tables, crops = make_workload(C, 1002 + rank)
    for j, t in enumerate(tables):
        pipe.upload_dict(j, t, n_bits=NBITS, ignore_bit=0, nonexist="zero")
    # Rotating input buffers: consecutive steps read DIFFERENT device buffers whose total exceeds the 126 MB L2 twice over,
    # so no step finds its logits in L2 (the timing rule's "inputs larger than L2"); buffer j is the batch rolled by j crops.
    L2_BYTES = 126 << 20
    n_buf = max(1, min(8, -(-2 * L2_BYTES // logits.nbytes) + 1)) if logits.nbytes < 2 * L2_BYTES else 1
    bufs = []
    for j in range(n_buf):
        r = (j * 7) % C
        bufs.append((torch.from_numpy(np.roll(logits, r, 0)).cuda(),
                     torch.from_numpy(np.roll(bboxes, r, 0).astype(np.float64)).cuda(),
                     torch.from_numpy(np.roll(Ks.reshape(C, 9), r, 0)).cuda(),
                     torch.from_numpy(np.roll(obj, r, 0).astype(np.int32)).cuda()))
    d_logits, d_bbox, d_K, d_obj = bufs[0]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    n_total = C * world
    kw = dict(n_bits=NBITS, iters=H, m=M, thr=THR)

    def final_gather(outs):
        """SURVEY 8(e) / configs[3]: the path has ONE exchange step, at the very end of the job -- the poses of all K
        steps of all ranks in one all_gather (112 B per crop), inside the timed region."""
        if world == 1:
            return outs
        poses = torch.cat([o[0] for o in outs]); ninl = torch.cat([o[1] for o in outs]); st = torch.cat([o[2] for o in outs])
        return zp.gather_poses(poses, ninl, st, poses.shape[0] * world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    outs = [pipe.submit(*bufs[i % n_buf], **kw) for i in range(max(args.warmup, args.lanes))]
    pipe.join()
    final_gather(outs)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = pipe.launch_count()
    barrier()
    # ---- timed region: exactly K steps, enqueued round-robin over the lanes, bracketed by barrier + synchronize
    e0.record()
    outs = [pipe.submit(*bufs[i % n_buf], **kw) for i in range(args.steps)]
    pipe.join()
    gathered = final_gather(outs)
    e1.record()
    barrier()
    launches = pipe.launch_count() - l0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = n_total * args.steps / (ms_total * 1e-3)

    # ---- latency of ONE step alone (single lane, L2 flushed by a 256 MiB write before it): reported beside the throughput
    lat = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.decode_and_pose_batch(d_logits, d_bbox, d_K, d_obj, **kw); b.record()
        b.synchronize()
        lat.append(a.elapsed_time(b))
    step_latency_ms = statistics.median(lat)

    # ---- e2e: HOST pinned buffers through the C-ABI host entry (zp_pose_batch_host_async on every lane + zp_sync): the
    # H2D copy of the logits and the D2H read of the poses are inside the timed region, every step
    h_logits = torch.from_numpy(logits).pin_memory()
    outs = [(torch.empty((C, 12), dtype=torch.float64).pin_memory().numpy(), torch.empty(C, dtype=torch.int32).pin_memory().numpy(),
             torch.empty(C, dtype=torch.int32).pin_memory().numpy()) for _ in range(args.lanes)]
    for i in range(max(3, args.lanes)):
        pipe.submit_host(h_logits, bboxes, Ks, obj, out=outs[pipe.next_lane])
    pipe.wait_host()
    barrier()
    e2e_steps = max(6, min(args.steps, 30))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        pipe.submit_host(h_logits, bboxes, Ks, obj, out=outs[pipe.next_lane])   # waits for (and so reads) that lane's previous result
    pipe.wait_host()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = n_total * e2e_steps / float(dt.item())
    h2d = logits.nbytes + C * 4 * 8 + C * 9 * 8 + C * 4
    d2h = C * 12 * 8 + C * 4 + C * 4
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    # ---- per-kernel timing (separate instrumented pass; CUDA events on the launching stream, L2 flushed)
    # the flush (512 MiB write, ~90 us of GPU time) also hides the CPU cost of enqueueing fn: the event pair and the
    # kernel are all queued before the flush retires, so the events bracket GPU time only
    flush2 = torch.empty(2 * L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        tot = 0.0
        for _ in range(reps):
            flush2.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            b.synchronize()
            tot += a.elapsed_time(b)
        return tot / reps

    def ktimed(fn, names, reps=20):
        """per-kernel durations from the library's own event pairs, recorded on the launching stream directly around each
        launch (zp_set_kernel_timing), L2 flushed before every repetition"""
        for _ in range(3):
            fn()
        eng.set_kernel_timing(True)
        for _ in range(reps):
            flush2.zero_()
            fn()
        torch.cuda.synchronize()
        out = {n: eng.kernel_time(n)[0] for n in names}
        eng.set_kernel_timing(False)
        return out

    line = None
    if rank == 0:
        corr, counts = eng.decode(d_logits, d_bbox, d_obj)
        cap = corr.shape[2]
        k_ms = ktimed(lambda: eng.decode(d_logits, d_bbox, d_obj), ["zp_decode_stream_kernel"])
        k_ms.update(ktimed(lambda: eng.ransac(corr, counts, d_K, H=H, m=M, thr=THR),
                           ["zp_samples_kernel", "zp_minimal_kernel", "zp_score_kernel", "zp_final_kernel"]))
        chain_event_ms = timed(lambda: eng.ransac(corr, counts, d_K, H=H, m=M, thr=THR))     # python-level events: + launch gaps, memset
        k_ms["ransac_chain(samples+minimal+score+select+final)"] = sum(k_ms[k] for k in ("zp_samples_kernel", "zp_minimal_kernel", "zp_score_kernel", "zp_final_kernel"))
        Mtot = int(counts.clamp(max=cap).sum().item())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["zp_decode_stream_kernel"].get(str(C))
        except Exception:
            pass
        dec_bytes = C * (1 + NBITS) * S * S * 4 + 20 * Mtot + 4 * C          # algorithmic HBM bytes
        dec_gbs = dec_bytes / (k_ms["zp_decode_stream_kernel"] * 1e-3) / 1e9
        fp32_scalar = eng.fp32_peak_tflops()
        fp32_packed = eng.fp32_peak_tflops(packed=True)
        fp32_peak = max(fp32_scalar, fp32_packed)
        sc_flops = 27.0 * H * Mtot                                            # 27 flop / (corr x hyp)
        sc_tf = sc_flops / (k_ms["zp_score_kernel"] * 1e-3) / 1e12
        chain = k_ms["zp_decode_stream_kernel"] + k_ms["ransac_chain(samples+minimal+score+select+final)"]
        shares = {k: round(v / chain, 4) for k, v in k_ms.items()}
        dominant = max(("zp_decode_stream_kernel", "zp_minimal_kernel", "zp_score_kernel", "zp_final_kernel"), key=lambda k: k_ms[k])
        if args.kernels:
            for k, v in k_ms.items():
                print("%-55s %9.3f us  share %.3f" % (k, v * 1e3, v / chain), file=sys.stderr)
        extras = None
        try:
            extras = next_rows(eng, timed, d_bbox, d_K, d_obj, crops, tables, C, hbm_peak, fp32_peak)
        except Exception as exc:          # the rows beside the path must never cost the headline line
            extras = {"error": repr(exc)[:200]}
        cpu = None
        if prev_affinity:
            os.sched_setaffinity(0, prev_affinity)          # the reference pool gets every host core again
        if not args.no_cpu_baseline:
            n_sample = max(256, 128 * (os.cpu_count() or 1))
            v, cores, secs, ref_poses = cpu_baseline(logits, bboxes, Ks, obj, tables, n_sample)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d crops of the same workload, fork pool of %d workers (cv2.setNumThreads(1)), %.1f s" % (n_sample, cores, secs)}
            try:
                cpu["agreement"] = agreement(eng, d_logits, d_bbox, d_K, d_obj, crops, tables, ref_poses, kw)
            except Exception as exc:
                cpu["agreement"] = {"error": repr(exc)[:200]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "collective": None if world == 1 else "one NCCL all_gather_into_tensor of the K steps' pose records (112 B/crop) at the end of the timed region",
            "dtype": "f32 scoring / f64 EPnP / u16 codes", "data": "synthetic",
            "config": {"workload": WORKLOAD % C,
                       "crops_per_gpu": C, "lanes": args.lanes,
                       "l2": "inputs larger than L2: %d rotating device batches of %.0f MB, a step never re-reads the buffer of "
                             "the previous %d steps (per-kernel figures: 512 MiB flush write before each launch)" % (n_buf, logits.nbytes / 1e6, n_buf - 1),
                       "masked_px_per_crop": Mtot / C},
            "step_latency_ms": step_latency_ms,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "path": "zp_pose_batch_host_async + zp_sync (C ABI, pinned host logits -> pinned host poses), %d lanes" % args.lanes,
                    "host_cpus": "%d CPUs local to the GPU (NVML affinity)" % len(numa_cpus) if numa_cpus else "unbound",
                    "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": {"kernel": "zp_decode_stream_kernel", "bound": "hbm", "achieved": dec_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": dec_gbs / hbm_peak, "traffic": traffic["bytes"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_launch": dec_bytes, "us_per_launch": k_ms["zp_decode_stream_kernel"] * 1e3},
            "roofline_score": {"kernel": "zp_score_kernel", "bound": "fp32", "achieved": sc_tf, "peak": fp32_peak,
                               "unit": "TFLOP/s", "frac": sc_tf / fp32_peak if fp32_peak else None,
                               "peak_source": "max of zp_fp32_peak_probe (scalar FFMA chains, %.1f) and zp_fp32x2_peak_probe (packed FFMA2 chains, %.1f), measured on this GPU in this run" % (fp32_scalar, fp32_packed),
                               "algorithmic_flops_per_launch": sc_flops, "us_per_launch": k_ms["zp_score_kernel"] * 1e3},
            "kernel_us": {k: round(v * 1e3, 2) for k, v in k_ms.items()},
            "kernel_us_method": "CUDA event pairs recorded by the library on the launching stream directly around each launch "
                                "(zp_set_kernel_timing), 512 MiB L2 flush before every repetition, 20 repetitions",
            "ransac_chain_us_python_events": round(chain_event_ms * 1e3, 2),
            "kernel_share_of_step": shares,
            "dominant_kernel": dominant,
            "cpu_baseline": cpu,
            "next_rows": extras,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
