"""Multi-GPU plumbing: samples are independent (the reference's only many-sample loop touches one sample per
iteration, src/evaluation_framework.py:534-556), so a batch is split into contiguous blocks, one per rank, each
rank runs the fused path on its shard with no data-path collective, and ONE all_gather moves the fixed-stride
per-sample result tables (boxes padded to `max_boxes`).  BEV grids stay sharded.  Works with NCCL (device
tensors) and gloo (host tensors, used by the CPU tests)."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the sample index range for `rank`: ceil(n/world) per rank, last ranks may be short."""
    per = (n_samples + world - 1) // world
    lo = min(rank * per, n_samples)
    return lo, min(lo + per, n_samples)


def pad_tables(host: Dict[str, np.ndarray], n_samples: int, n_cams: int, max_boxes: int, per_rank: int) -> Dict[str, torch.Tensor]:
    """Fixed-stride tables [per_rank, max_boxes, ...] + n_boxes column, from the ragged per-box arrays of one shard."""
    off = host["sample_box_off"]
    out = {
        "n_boxes": torch.zeros(per_rank, dtype=torch.int32),
        "box_count": torch.zeros((per_rank, max_boxes), dtype=torch.int32),
        "box_nearest": torch.full((per_rank, max_boxes), float("inf"), dtype=torch.float32),
        "box_centroid": torch.zeros((per_rank, max_boxes, 3), dtype=torch.float32),
        "proj_visible": torch.zeros((per_rank, max_boxes, n_cams), dtype=torch.uint8),
        "proj_extent": torch.zeros((per_rank, max_boxes, n_cams, 4), dtype=torch.float32),
        "stats": torch.zeros((per_rank, 16), dtype=torch.int32),
    }
    for i in range(n_samples):
        b0, b1 = int(off[i]), int(off[i + 1])
        nb = b1 - b0
        out["n_boxes"][i] = nb
        out["box_count"][i, :nb] = torch.from_numpy(host["box_count"][b0:b1].view(np.int32))
        out["box_nearest"][i, :nb] = torch.from_numpy(host["box_nearest"][b0:b1])
        out["box_centroid"][i, :nb] = torch.from_numpy(host["box_centroid"][b0:b1])
        out["proj_visible"][i, :nb] = torch.from_numpy(host["proj_visible"][b0:b1])
        out["proj_extent"][i, :nb] = torch.from_numpy(host["proj_extent"][b0:b1])
        out["stats"][i] = torch.from_numpy(host["stats"][i].view(np.int32))
    return out


def gather_tables(tables: Dict[str, torch.Tensor], device=None) -> Dict[str, torch.Tensor]:
    """all_gather of every table along a new leading rank axis; returns [world * per_rank, ...] tensors on every rank."""
    world = dist.get_world_size()
    out = {}
    for k, t in tables.items():
        t = t.contiguous() if device is None else t.to(device).contiguous()
        buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)  # rank-major concat
        dist.all_gather_into_tensor(buf, t)
        out[k] = buf
    return out
