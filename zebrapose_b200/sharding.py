"""Multi-GPU: crops are independent, so they shard as contiguous ranges over ranks with no data-path collective; one
all_gather of the fixed-size pose records at the end.  Host logic only (torch.distributed:
NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist

RECORD = 14   # 12 pose + n_inliers + status


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of crops owned by `rank`: ceil(n/world) per rank, the tail ranks may be short/empty."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def pack_records(poses, n_inliers, status):
    """[n_local,12] f64, [n_local] i32, [n_local] i32 -> [n_local,14] f64 records."""
    return torch.cat([poses.reshape(-1, 12).to(torch.float64), n_inliers.reshape(-1, 1).to(torch.float64),
                      status.reshape(-1, 1).to(torch.float64)], 1)


def gather_poses(poses, n_inliers, status, n_total, group=None):
    """Final gather of the per-rank results: every rank returns (poses [n_total,12] f64, n_inliers [n_total] i32,
    status [n_total] i32) in global crop order.  Shards are padded to ceil(n/world) records so one
    all_gather_into_tensor of 112 B/crop suffices."""
    if not (dist.is_available() and dist.is_initialized()):
        return poses, n_inliers, status
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_total + world - 1) // world
    rec = pack_records(poses, n_inliers, status)
    lo, hi = shard_range(n_total, rank, world)
    if rec.shape[0] != hi - lo:
        raise ValueError("rank %d holds %d records, its shard is [%d,%d)" % (rank, rec.shape[0], lo, hi))
    buf = torch.zeros((per, RECORD), dtype=torch.float64, device=rec.device)
    buf[: rec.shape[0]] = rec
    out = torch.empty((world * per, RECORD), dtype=torch.float64, device=rec.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    # global crop i lives at row (i // per) * per + i % per = i, so the padding of short tail shards is at the end
    out = out[:n_total]
    return out[:, :12].contiguous(), out[:, 12].to(torch.int32), out[:, 13].to(torch.int32)
