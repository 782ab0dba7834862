"""ctypes bridge to the CPU oracle (oracle/libmsc_oracle.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libmsc_oracle.so")
_lib = None


class OrcParams(C.Structure):
    _fields_ = [("remove_close_radius", C.c_float), ("range_min", C.c_float), ("range_max", C.c_float), ("z_min", C.c_float),
                ("z_max", C.c_float), ("ground_z", C.c_float), ("bev_range", C.c_float), ("bev_res", C.c_int32),
                ("image_w", C.c_int32), ("image_h", C.c_int32), ("n_cams", C.c_int32), ("fov_keep_mask", C.c_uint32),
                ("centroid_shift", C.c_int32), ("intensity_shift", C.c_int32)]


def build():
    subprocess.run(["make", "-C", os.path.join(_ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_aggregate_sweeps.restype = C.c_uint32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def orc_params(p) -> OrcParams:
    return OrcParams(p.remove_close_radius, p.range_min, p.range_max, p.z_min, p.z_max, p.ground_z, p.bev_range, p.bev_res,
                     p.image_w, p.image_h, p.n_cams, p.fov_keep_mask, p.centroid_shift, p.intensity_shift)


def oracle_fused(hb, i: int, p) -> dict:
    """Run the scalar definition of the fused path on sample i of a HostBatch."""
    L = lib()
    s0, s1 = int(hb.sample_sweep_off[i]), int(hb.sample_sweep_off[i + 1])
    b0, b1 = int(hb.sample_box_off[i]), int(hb.sample_box_off[i + 1])
    nb, nc, R = b1 - b0, p.n_cams, p.bev_res
    start = np.ascontiguousarray(hb.sweep_start[s0:s1]); count = np.ascontiguousarray(hb.sweep_count[s0:s1])
    pose = np.ascontiguousarray(hb.sweep_pose[s0:s1]); boxes = np.ascontiguousarray(hb.boxes[b0:b1])
    ego = np.ascontiguousarray(hb.ego_pose[i]); lcal = np.ascontiguousarray(hb.lidar_calib[i])
    cpose = np.ascontiguousarray(hb.cam_ego_pose[i]); ccal = np.ascontiguousarray(hb.cam_calib[i]); cK = np.ascontiguousarray(hb.cam_K[i])
    out = {"box_count": np.zeros(nb, np.uint32), "box_nearest": np.zeros(nb, np.float32), "box_centroid": np.zeros((nb, 3), np.float32),
           "bev_count": np.zeros((R, R), np.uint32), "bev_height": np.zeros((R, R), np.float32), "bev_isum_q": np.zeros((R, R), np.uint32),
           "stats": np.zeros(16, np.uint32), "proj_visible": np.zeros((nb, nc), np.uint8), "proj_extent": np.zeros((nb, nc, 4), np.float32)}
    op = orc_params(p)
    L.orc_fused_evidence_sample(C.byref(op), _p(hb.points), C.c_int(s1 - s0), _p(start), _p(count), _p(pose), C.c_int(nb), _p(boxes),
                                _p(ego), _p(lcal), _p(ccal), _p(cK), _p(out["box_count"]), _p(out["box_nearest"]), _p(out["box_centroid"]),
                                _p(out["bev_count"]), _p(out["bev_height"]), _p(out["bev_isum_q"]), _p(out["stats"]))
    L.orc_project_boxes(C.c_int(nb), _p(boxes), C.c_int(nc), _p(cpose), _p(ccal), _p(cK), C.c_int(p.image_w), C.c_int(p.image_h),
                        _p(out["proj_visible"]), _p(out["proj_extent"]))
    return out


def oracle_aggregate(hb, i: int, remove_close_radius: float = 1.0):
    L = lib()
    s0, s1 = int(hb.sample_sweep_off[i]), int(hb.sample_sweep_off[i + 1])
    start = np.ascontiguousarray(hb.sweep_start[s0:s1]); count = np.ascontiguousarray(hb.sweep_count[s0:s1])
    pose = np.ascontiguousarray(hb.sweep_pose[s0:s1]); lag = np.ascontiguousarray(hb.sweep_time_lag[s0:s1])
    n = int(count.sum())
    xyzi = np.zeros((n, 4), np.float32); tl = np.zeros(n, np.float32)
    m = L.orc_aggregate_sweeps(C.c_float(remove_close_radius), _p(hb.points), C.c_int(s1 - s0), _p(start), _p(count), _p(pose), _p(lag),
                               _p(xyzi), _p(tl))
    return xyzi[:m].copy(), tl[:m].copy()


def oracle_project(boxes, cam_pose, cam_calib, cam_K, W=1600, H=900):
    L = lib()
    boxes = np.ascontiguousarray(boxes, np.float64); nb = boxes.shape[0]; nc = cam_pose.shape[0]
    vis = np.zeros((nb, nc), np.uint8); ext = np.zeros((nb, nc, 4), np.float32)
    L.orc_project_boxes(C.c_int(nb), _p(boxes), C.c_int(nc), _p(np.ascontiguousarray(cam_pose)), _p(np.ascontiguousarray(cam_calib)),
                        _p(np.ascontiguousarray(cam_K)), C.c_int(W), C.c_int(H), _p(vis), _p(ext))
    return vis, ext


def oracle_annotation_table(xy, vel):
    L = lib()
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2); vel = np.ascontiguousarray(vel, np.float64).reshape(-1, 2)
    n = xy.shape[0]
    out = {"distance": np.zeros(n), "direction": np.zeros(n, np.uint8), "moving": np.zeros(n, np.uint8), "zone": np.zeros(n, np.uint8),
           "region_bits": np.zeros(n, np.uint8)}
    L.orc_annotation_table(C.c_int(n), _p(xy), _p(vel), _p(out["distance"]), _p(out["direction"]), _p(out["moving"]), _p(out["zone"]),
                           _p(out["region_bits"]))
    return out


def oracle_footprints(boxes, ego_pose=None):
    L = lib()
    boxes = np.ascontiguousarray(boxes, np.float64); n = boxes.shape[0]
    rect = np.zeros((n, 6))
    ego = None if ego_pose is None else np.ascontiguousarray(ego_pose, np.float64)
    L.orc_box_footprints(C.c_int(n), _p(boxes), _p(ego) if ego is not None else None, _p(rect))
    return rect


def oracle_relations(rect):
    L = lib()
    rect = np.ascontiguousarray(rect, np.float64); n = rect.shape[0]
    out = {"dist": np.zeros((n, n), np.float32), "bearing": np.zeros((n, n), np.float32), "category": np.zeros((n, n), np.uint8),
           "overlap": np.zeros((n, n), np.uint8)}
    L.orc_relation_table(C.c_int(n), _p(rect), _p(out["dist"]), _p(out["bearing"]), _p(out["category"]), _p(out["overlap"]))
    return out


def oracle_keyframe_filter_split(pts, p):
    """pts: (N, >=4) float32 array, possibly a strided view (row pitch in floats is derived from strides)."""
    L = lib()
    pts = np.asarray(pts)
    assert pts.dtype == np.float32 and pts.strides[1] == 4
    pitch = pts.strides[0] // 4
    n = pts.shape[0]
    kept = np.zeros(n, np.uint32); g = np.zeros(n, np.uint32); o = np.zeros(n, np.uint32)
    nk, ng, no = C.c_uint32(), C.c_uint32(), C.c_uint32()
    L.orc_keyframe_filter_split(C.c_void_p(pts.ctypes.data), C.c_uint32(n), C.c_int(pitch), C.c_float(p.range_min), C.c_float(p.range_max),
                                C.c_float(p.z_min), C.c_float(p.z_max), C.c_float(p.ground_z), _p(kept), C.byref(nk), _p(g), C.byref(ng),
                                _p(o), C.byref(no))
    return kept[:nk.value].copy(), g[:ng.value].copy(), o[:no.value].copy()


def oracle_keyframe_bev(pts, ground_idx, object_idx, bev_range=50.0, res=800):
    L = lib()
    pts = np.asarray(pts)
    assert pts.dtype == np.float32 and pts.strides[1] == 4
    pitch = pts.strides[0] // 4
    count = np.zeros((res, res), np.uint32); height = np.zeros((res, res), np.float32); sem = np.zeros((res, res, 3), np.uint8)
    gi = np.ascontiguousarray(ground_idx, np.uint32); oi = np.ascontiguousarray(object_idx, np.uint32)
    L.orc_keyframe_bev_raster(C.c_void_p(pts.ctypes.data), C.c_int(pitch), _p(gi), C.c_uint32(gi.size), _p(oi), C.c_uint32(oi.size),
                              C.c_float(bev_range), C.c_int(res), _p(count), _p(height), _p(sem))
    return count, height, sem


def oracle_cloud_stats(pts):
    L = lib()
    pts = np.asarray(pts); pitch = pts.strides[0] // 4
    mm = np.zeros(6, np.float32); acc = C.c_double()
    L.orc_cloud_stats(C.c_void_p(pts.ctypes.data), C.c_uint32(pts.shape[0]), C.c_int(pitch), _p(mm), C.byref(acc))
    return mm, acc.value


def oracle_cluster_aabb(pts, labels, n_clusters):
    L = lib()
    pts = np.asarray(pts); pitch = pts.strides[0] // 4
    labels = np.ascontiguousarray(labels, np.int32)
    out = np.zeros((n_clusters, 11), np.float32)
    L.orc_cluster_aabb(C.c_void_p(pts.ctypes.data), C.c_uint32(pts.shape[0]), C.c_int(pitch), _p(labels), C.c_int(n_clusters), _p(out))
    return out
