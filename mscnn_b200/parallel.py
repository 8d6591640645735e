"""Image-parallel execution across the GPUs of one box.

Images are independent, so a global batch is split into contiguous per-rank shards with no
data-path collective; the only exchange is ONE all-gather of the final detections (fixed-size
padded [B, cap, 5] boxes + [B] counts per rank) at the end of a step.  The reference has no
multi-GPU inference at all (its P2PSync is training-only, src/caffe/parallel.cpp:421-439).

torch.distributed is plumbing here: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


def shard_range(rank: int, per_rank: int) -> tuple[int, int]:
    """Global image indices [first, last) of a rank (weak scaling: `per_rank` images each)."""
    return rank * per_rank, (rank + 1) * per_rank


@dataclass
class GatherBuffers:
    dets: torch.Tensor     # [world, B, cap, 5]
    counts: torch.Tensor   # [world, B]

    def __init__(self, world: int, batch: int, cap: int, device):
        self.dets = torch.zeros((world, batch, cap, 5), dtype=torch.float32, device=device)
        self.counts = torch.zeros((world, batch), dtype=torch.int32, device=device)


def all_gather_detections(dets: torch.Tensor, counts: torch.Tensor, buf: GatherBuffers) -> None:
    """dets [B, cap, 5] / counts [B] of this rank -> buf on every rank (rank-major = global image order)."""
    if dist.get_backend() == "gloo":   # gloo has no all_gather_into_tensor for all dtypes
        dl = list(buf.dets.unbind(0))
        cl = list(buf.counts.unbind(0))
        dist.all_gather(dl, dets.contiguous())
        dist.all_gather(cl, counts.contiguous())
        return
    dist.all_gather_into_tensor(buf.dets, dets)
    dist.all_gather_into_tensor(buf.counts, counts)


def merge_detections(buf: GatherBuffers) -> list[torch.Tensor]:
    """Per global image (rank-major order) the [K, 5] = [x y w h prob] detections, padding stripped."""
    world, b = buf.counts.shape
    out = []
    for r in range(world):
        for i in range(b):
            out.append(buf.dets[r, i, : int(buf.counts[r, i])].clone())
    return out
