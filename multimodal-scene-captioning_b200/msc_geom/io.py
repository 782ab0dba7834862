"""On-disk format step: nuScenes `.pcd.bin` sweeps -> the fused kernel's batch layout.

The reference reads one keyframe per sample through the devkit (`LidarPointCloud.from_file`: np.fromfile -> reshape(-1,5)[:, :4].T,
then `.T` again, src/nuscenes_loader.py:146-157) and its evaluator loads every sample of every scene once just to collect tokens
(src/evaluation_framework.py:488-496).  Here the raw 5-float rows are read with `readinto` straight into their final place in one
staging buffer (pinned when a CUDA device is present) -- no intermediate arrays, no repack -- by a small thread pool (file reads
release the GIL), and scenes can be scanned for tokens without touching sensor data.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence

import numpy as np

from .layout import HostBatch, N_CAMS_DEFAULT, boxes_from_annotations

ROW_BYTES = 20  # x, y, z, intensity, ring as float32


def pcd_bin_points(path: str) -> int:
    """Number of points in a .pcd.bin file (size / 20 bytes); raises on a truncated file."""
    size = os.path.getsize(path)
    if size % ROW_BYTES:
        raise ValueError(f"{path}: size {size} is not a multiple of {ROW_BYTES} bytes")
    return size // ROW_BYTES


def read_pcd_bin_into(path: str, dst_rows: np.ndarray) -> None:
    """Read a .pcd.bin file into `dst_rows` ((n,5) float32, C-contiguous view of the staging buffer) without a temporary."""
    buf = memoryview(dst_rows).cast("B")
    with open(path, "rb", buffering=0) as f:
        got = 0
        while got < len(buf):
            k = f.readinto(buf[got:])
            if not k:
                raise IOError(f"{path}: short read ({got} of {len(buf)} bytes)")
            got += k


def stage_batch(samples: Sequence[dict], n_cams: int = N_CAMS_DEFAULT, threads: int = 8, pinned: Optional[bool] = None) -> HostBatch:
    """Build a HostBatch whose sweeps come from files.  Each sample is a dict shaped like the loader's, where a sweep is
    {"path": <.pcd.bin>, "ref_from_sensor": 3x4 f64, "time_lag": s} (or carries "points_raw" already in memory).  The `points`
    buffer is allocated once (pinned host memory when CUDA is available, unless `pinned=False`) and files are read into it in place."""
    S = len(samples)
    sweeps = [(i, sw) for i, s in enumerate(samples) for sw in s["lidar_sweeps"]]
    counts = np.array([pcd_bin_points(sw["path"]) if "path" in sw else int(np.asarray(sw["points_raw"]).shape[0]) for _, sw in sweeps], np.uint32)
    starts = np.zeros(len(sweeps), np.uint32)
    cursor = 0
    for k, c in enumerate(counts):
        starts[k] = cursor
        cursor += (int(c) + 3) & ~3
    n_rows = cursor + 4
    if pinned is None:
        try:
            import torch
            pinned = torch.cuda.is_available()
        except Exception:
            pinned = False
    if pinned:
        import torch
        holder = torch.empty((n_rows, 5), dtype=torch.float32).pin_memory()
        points = holder.numpy()
    else:
        holder = None
        points = np.empty((n_rows, 5), np.float32)
    # padding rows (between sweeps and at the tail) must fail every compare
    for k in range(len(sweeps)):
        a, c = int(starts[k]) + int(counts[k]), (int(starts[k + 1]) if k + 1 < len(sweeps) else cursor)
        points[a:c] = np.nan
    points[cursor:] = np.nan

    def load(k):
        sw = sweeps[k][1]
        dst = points[int(starts[k]): int(starts[k]) + int(counts[k])]
        if "path" in sw:
            read_pcd_bin_into(sw["path"], dst)
        else:
            dst[:] = sw["points_raw"]

    with ThreadPoolExecutor(max(1, threads)) as ex:
        list(ex.map(load, range(len(sweeps))))
    sample_sweep_off = np.zeros(S + 1, np.int32)
    for i, _ in sweeps:
        sample_sweep_off[i + 1] += 1
    sample_sweep_off = np.cumsum(sample_sweep_off).astype(np.int32)
    pose = np.stack([np.asarray(sw["ref_from_sensor"], np.float64).reshape(-1)[:12] for _, sw in sweeps]) if sweeps else np.zeros((0, 12))
    lag = np.array([float(sw.get("time_lag", 0.0)) for _, sw in sweeps], np.float32)
    box_arrays = [boxes_from_annotations(s.get("annotations", [])) for s in samples]
    sample_box_off = np.zeros(S + 1, np.int32)
    sample_box_off[1:] = np.cumsum([b.shape[0] for b in box_arrays])
    boxes = np.ascontiguousarray(np.concatenate(box_arrays, 0) if box_arrays else np.zeros((0, 10)), np.float64).reshape(-1, 10)
    ident = np.array([0, 0, 0, 1, 0, 0, 0.0])
    ego = np.stack([np.asarray(s.get("ego_pose", ident), np.float64) for s in samples]) if S else np.zeros((0, 7))
    lcal = np.stack([np.asarray(s.get("lidar_calib", ident), np.float64) for s in samples]) if S else np.zeros((0, 7))
    cam_pose = np.zeros((S, n_cams, 7)); cam_cal = np.zeros((S, n_cams, 7)); cam_K = np.zeros((S, n_cams, 9))
    cam_pose[..., 3] = 1.0; cam_cal[..., 3] = 1.0
    cam_K[..., 0] = cam_K[..., 4] = cam_K[..., 8] = 1.0
    for i, s in enumerate(samples):
        for c, cam in enumerate(s.get("cameras", [])[:n_cams]):
            cam_pose[i, c] = cam.get("ego_pose", ego[i]); cam_cal[i, c] = cam["calib"]; cam_K[i, c] = np.asarray(cam["intrinsic"], np.float64).reshape(-1)
    hb = HostBatch(S, points, sample_sweep_off, starts, counts, np.ascontiguousarray(pose, np.float64), lag, sample_box_off, boxes, ego, lcal,
                   cam_pose, cam_cal, cam_K, n_cams, int(max([b.shape[0] for b in box_arrays], default=0)))
    hb._pinned_holder = holder  # keeps the pinned allocation alive with the batch
    return hb


def write_pcd_bin(path: str, rows: np.ndarray) -> None:
    """Write (n,5) float32 rows in the nuScenes .pcd.bin layout (used by tests and synthetic datasets)."""
    np.ascontiguousarray(rows, np.float32).tofile(path)
