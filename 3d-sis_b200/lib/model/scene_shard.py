"""Scene-sharded multi-GPU inference: every .scene/.chunk is an independent unit
(batch size 1, no cross-scene state; reference: lib/model/trainval.py:787-822), so rank r of W
processes its share of the scene list with NO collective on the data path.  torch.distributed
(NCCL on the GPUs, gloo in the CPU tests) only carries (a) the final gather of the small,
variable-length detection tensors and (b) the max-over-ranks time for the throughput figure.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_scenes(costs, rank, world):
    """Longest-processing-time-first assignment of scenes to ranks by cost (voxel count).
    Returns the scene indices of `rank`, in processing order.  Deterministic on every rank."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += float(costs[i])
        if r == rank:
            mine.append(i)
    return mine


def gather_detections(local, device=None, group=None):
    """local: list of (scene_index, boxes[k,6], classes[k], conf[k]) tensors of this rank.
    Returns on every rank a dict scene_index -> (boxes, classes, conf) for all scenes: one all_gather of
    the counts and one padded all_gather of a [rows, 9] float tensor (scene, 6 box, class, conf)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rows = []
    for idx, boxes, cls, conf in local:
        k = boxes.shape[0]
        marker = torch.zeros(1, 9)
        marker[0, 0], marker[0, 7] = float(idx), -1.0  # class -1: "scene processed" marker (scenes may have 0 boxes)
        rows.append(marker)
        if k:
            rows.append(torch.cat([torch.full((k, 1), float(idx)), boxes.detach().float().cpu().reshape(k, 6),
                                   cls.detach().float().cpu().reshape(k, 1), conf.detach().float().cpu().reshape(k, 1)], 1))
    mine = torch.cat(rows, 0) if rows else torch.zeros(0, 9)
    if world == 1:
        allrows = [mine]
    else:
        device = device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
        n = torch.tensor([mine.shape[0]], dtype=torch.int64, device=device)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n, group=group)
        cap = max(int(c.item()) for c in counts)
        pad = torch.zeros(max(cap, 1), 9, device=device)
        pad[:mine.shape[0]] = mine.to(device)
        bufs = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        allrows = [b[:int(c.item())].cpu() for b, c in zip(bufs, counts)]
    out = {}
    for t in allrows:
        for idx in t[:, 0].unique().tolist():
            sel = t[(t[:, 0] == idx) & (t[:, 7] >= 0)]
            out[int(idx)] = (sel[:, 1:7], sel[:, 7].long(), sel[:, 8])
    return out


def max_over_ranks(value, device=None, group=None):
    """Device-timed milliseconds -> max over ranks (the contract's timing rule)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    device = device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
