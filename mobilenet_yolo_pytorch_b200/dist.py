"""Data-parallel sharding of the hot path over the GPUs of one box (one process
per GPU, torch.distributed; NCCL over NVLink on the B200s, gloo in CPU tests).

The batch shards by image (images are independent in decode + NMS,
utils/box.py:16, yolo_loss.py:202-203).  The only exchange steps are
  * an all-gather of the fixed-stride detections + counts, and
  * an all-reduce(SUM) of the 16 loss partial sums (the normalisers of
    yolo_loss.py:55,224,170-178 are batch-global).
No collective sits inside the data path."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced chunks: the first n % world ranks get one extra image."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_detections(dets: torch.Tensor, counts: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """dets (n_local, K, 7) fp32 and counts (n_local,) int32 of this rank's shard ->
    the same for the whole batch in rank order.  Shards must have equal n_local
    (pad the batch to a multiple of the world size).  The counts travel inside the
    same message as one extra row per image, so this is ONE collective."""
    world = dist.get_world_size(group)
    n, K, _ = dets.shape
    packed = torch.empty((n, K + 1, 7), dtype=torch.float32, device=dets.device)
    packed[:, :K] = dets
    packed[:, K, 0] = counts.to(torch.float32)  # exact for counts < 2^24
    out = torch.empty((world * n, K + 1, 7), dtype=torch.float32, device=dets.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return out[:, :K], out[:, K, 0].to(torch.int32)


def all_gather_detections_compact(dets: torch.Tensor, counts: torch.Tensor, group=None):
    """Like ``all_gather_detections`` but only kept rows travel: the shard's rows are packed back to back on the device
    (``b200yolo_compact_rows``), the per-rank totals and per-image counts are gathered (tiny), and ONE all-gather moves
    ``max total`` rows per rank -- 0.5 MB instead of 13 MB per rank for trained-like heads at 256 images.
    Returns (rows (sum of totals, 7) in rank-then-image order, counts (world*n_local,) int32)."""
    from . import ops
    world = dist.get_world_size(group)
    n = dets.shape[0]
    packed, offsets = ops.compact_rows(dets, counts)
    totals = torch.empty((world,), dtype=torch.int32, device=dets.device)
    dist.all_gather_into_tensor(totals, offsets[n:n + 1].contiguous(), group=group)
    all_counts = torch.empty((world * n,), dtype=torch.int32, device=dets.device)
    dist.all_gather_into_tensor(all_counts, counts.contiguous(), group=group)
    tot = totals.cpu().tolist()                      # the one host sync: buffer sizes
    m = max(max(tot), 1)
    recv = torch.empty((world, m, 7), dtype=torch.float32, device=dets.device)
    dist.all_gather_into_tensor(recv, packed[:m].contiguous(), group=group)
    rows = torch.cat([recv[r, :tot[r]] for r in range(world)], 0)
    return rows, all_counts


def all_reduce_loss_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def split_targets(targets: List, world_size: int, rank: int) -> List:
    lo, hi = shard_bounds(len(targets), world_size, rank)
    return targets[lo:hi]
