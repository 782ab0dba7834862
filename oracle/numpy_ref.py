"""numpy_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of the functions of AgustinRoca/multimodal-scene-captioning that sit on the hot path
(SURVEY.md section 8(a), rows a3-a13).  Every function cites the reference lines it follows and is pinned
against golden vectors produced by importing the reference itself (tests/golden/make_golden.py); see
tests/test_oracle_golden.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import it.

Scalar promotion follows NumPy 2 (NEP 50), which is what the reference executes under in this environment
(SURVEY.md section 0.4): float32 arrays stay float32 against Python scalars.
"""
from __future__ import annotations

from collections import Counter
from typing import Dict, List, Optional, Tuple

import numpy as np

DIRECTIONS_8 = ["front_right", "front", "front_left", "left", "back_left", "back", "back_right", "right"]
DIRECTIONS_4 = ["front", "left", "back", "right"]
# scenegraph_agent.py:136-146, in dict order
SPATIAL_ZONES = [("front_close", 0, 10, "front"), ("front_medium", 10, 30, "front"), ("front_far", 30, 50, "front"),
                 ("left_close", 0, 10, "left"), ("left_medium", 10, 30, "left"), ("right_close", 0, 10, "right"),
                 ("right_medium", 10, 30, "right"), ("back_close", 0, 10, "back"), ("back_medium", 10, 30, "back")]


def preprocess_point_cloud(pc: np.ndarray, bev_range=50) -> np.ndarray:
    """lidar_agent.py:103-112"""
    distances = np.sqrt(pc[:, 0] ** 2 + pc[:, 1] ** 2)
    valid = (distances > 1.0) & (distances < bev_range)
    valid &= (pc[:, 2] < 5.0) & (pc[:, 2] > -3.0)
    return pc[valid]


def segment_ground(pc: np.ndarray, ground_threshold: float = -1.4) -> Tuple[np.ndarray, np.ndarray]:
    """lidar_agent.py:114-132"""
    m = pc[:, 2] < ground_threshold
    return pc[m], pc[~m]


def get_direction(position_2d) -> str:
    """lidar_agent.py:506-530 (labels are rotated 45 degrees from geometry in the reference; preserved)."""
    x, y = position_2d
    angle = np.arctan2(y, x) * 180 / np.pi
    angle = (angle + 360) % 360
    if 337.5 <= angle or angle < 22.5:
        return "front_right"
    for k, name in enumerate(DIRECTIONS_8[1:], start=1):
        if 22.5 + 45.0 * (k - 1) <= angle < 22.5 + 45.0 * k:
            return name
    return "right"


def to_pixels(coords: np.ndarray, r, res):
    """lidar_agent.py:547-552"""
    x_pix = ((coords[:, 0] + r) / (2 * r) * res).astype(int)
    y_pix = ((coords[:, 1] + r) / (2 * r) * res).astype(int)
    return np.clip(x_pix, 0, res - 1), np.clip(y_pix, 0, res - 1)


def bev_raster_layers(ground: np.ndarray, obj: np.ndarray, res=800, r=50):
    """The raster half of _generate_multi_layer_bev, lidar_agent.py:539-597, vectorised but order-faithful:
    returns (count int64, height f32 0-initialised running max, semantic BGR u8) BEFORE normalisation, the ego
    cross, the flips and the overlays."""
    height = np.zeros((res, res), np.float32)
    count = np.zeros((res, res), np.int64)
    allp = np.vstack([ground, obj])
    x, y = to_pixels(allp, r, res)
    np.add.at(count, (y, x), 1)
    np.maximum.at(height, (y, x), allp[:, 2].astype(np.float32))  # :560 max(height, z) starting from 0
    sem = np.zeros((res, res, 3), np.uint8)
    gx, gy = to_pixels(ground, r, res)
    sem[gy, gx] = [80, 80, 120]
    ox, oy = to_pixels(obj, r, res)
    h = obj[:, 2]
    if len(h) > 0 and h.max() > h.min():
        norm = (h - h.min()) / (h.max() - h.min())
    else:
        norm = np.ones(len(h)) * 0.5
    # last writer wins (:584-597): process in array order
    for xx, yy, hn in zip(ox, oy, norm):
        if hn < 0.5:
            g = int(255 * (1 - hn * 2))
        else:
            g = int(255 * (1 - (hn - 0.5) * 2))
        sem[yy, xx] = [0, g, 255]
    return count, height, sem


def density_from_count(count: np.ndarray) -> np.ndarray:
    """lidar_agent.py:563-564 (float32 log1p, divide by max, * 255, truncating uint8 cast)."""
    d = np.log1p(count.astype(np.float32))
    return (d / d.max() * 255).astype(np.uint8) if d.max() > 0 else d.astype(np.uint8)


def finish_bev(count, height, sem, res=800, r=50) -> Dict[str, np.ndarray]:
    """lidar_agent.py:563-564 and :599-642: density normalisation, ego cross, vertical flips, range rings, labels."""
    import cv2
    density = density_from_count(count)
    vis = np.ascontiguousarray(sem)
    center = res // 2
    ms = 15
    cv2.line(vis, (center - ms, center), (center + ms, center), (0, 255, 0), 3)
    cv2.line(vis, (center, center - ms), (center, center + ms), (0, 255, 0), 3)
    vis = cv2.flip(vis, 0)
    height = cv2.flip(np.ascontiguousarray(height), 0)
    density = cv2.flip(np.ascontiguousarray(density), 0)
    for dist in [10, 20, 30, 40]:
        radius = int(dist / (2 * r) * res)
        cv2.circle(vis, (center, center), radius, (100, 100, 100), 1)
        cv2.putText(vis, f"{dist}m", (center + 5, center - radius + 15), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (150, 150, 150), 1)
    cv2.putText(vis, "FRONT", (center - 25, 20), cv2.FONT_HERSHEY_SIMPLEX, 0.6, (200, 200, 200), 2)
    cv2.putText(vis, "BACK", (center - 20, res - 10), cv2.FONT_HERSHEY_SIMPLEX, 0.6, (200, 200, 200), 2)
    cv2.putText(vis, "L", (10, center + 5), cv2.FONT_HERSHEY_SIMPLEX, 0.6, (200, 200, 200), 2)
    cv2.putText(vis, "R", (res - 20, center + 5), cv2.FONT_HERSHEY_SIMPLEX, 0.6, (200, 200, 200), 2)
    return {"semantic": vis, "height": height, "density": density}


def generate_multi_layer_bev(ground, obj, res=800, r=50) -> Dict[str, np.ndarray]:
    """lidar_agent.py:532-642"""
    return finish_bev(*bev_raster_layers(ground, obj, res, r), res=res, r=r)


def cluster_metadata(cluster_points: np.ndarray) -> dict:
    """lidar_agent.py:200-218"""
    mn = cluster_points[:, :3].min(axis=0)
    mx = cluster_points[:, :3].max(axis=0)
    center = (mn + mx) / 2
    return {"center": center, "dimensions": mx - mn, "distance": np.sqrt(center[0] ** 2 + center[1] ** 2),
            "direction": get_direction(center[:2]), "num_points": len(cluster_points)}


def parse_annotations(annotations: List[dict]) -> List[dict]:
    """scenegraph_agent.py:180-247"""
    out = []
    for i, ann in enumerate(annotations):
        pos = ann.get("translation", [0, 0, 0])
        distance = np.sqrt(pos[0] ** 2 + pos[1] ** 2)
        angle = np.arctan2(pos[1], pos[0]) * 180 / np.pi
        angle = (angle + 360) % 360
        if 45 <= angle < 135:
            direction = "front"
        elif 135 <= angle < 225:
            direction = "left"
        elif 225 <= angle < 315:
            direction = "back"
        else:
            direction = "right"
        category = ann.get("category_name", "unknown").lower()
        for prefix in ["vehicle.", "human.pedestrian.", "movable_object.", "static_object."]:
            category = category.replace(prefix, "")
        velocity = ann.get("velocity", None)
        state = "stopped"
        if velocity is not None:
            try:
                if isinstance(velocity, (list, tuple)) and len(velocity) >= 2:
                    vx, vy = velocity[0], velocity[1]
                    if vx is not None and vy is not None:
                        state = "moving" if np.sqrt(vx ** 2 + vy ** 2) > 0.5 else "stopped"
            except (TypeError, IndexError, ValueError):
                state = "stopped"
        vis = str(ann.get("visibility_token", ""))
        if "80" in vis or "100" in vis:
            visibility = "high"
        elif "40" in vis or "60" in vis:
            visibility = "medium"
        else:
            visibility = "low"
        out.append({"id": f"obj_{i}", "category": category, "position": pos, "distance": distance, "direction": direction,
                    "state": state, "visibility": visibility, "attributes": ann.get("attribute_tokens", [])})
    return out


def categorize_objects(objects: List[dict]) -> Dict[str, List[dict]]:
    """scenegraph_agent.py:249-279"""
    cats = {k: [] for k in ["vehicles", "cyclists", "pedestrians", "barriers", "traffic_cones", "construction", "other"]}
    for obj in objects:
        c = obj["category"]
        if "car" in c or "truck" in c or "bus" in c or "trailer" in c:
            cats["vehicles"].append(obj)
        elif "bicycle" in c or "motorcycle" in c:
            cats["cyclists"].append(obj)
        elif "pedestrian" in c or "adult" in c or "child" in c:
            cats["pedestrians"].append(obj)
        elif "barrier" in c:
            cats["barriers"].append(obj)
        elif "cone" in c:
            cats["traffic_cones"].append(obj)
        elif "construction" in c:
            cats["construction"].append(obj)
        else:
            cats["other"].append(obj)
    return cats


def build_spatial_zones(objects: List[dict]) -> Dict[str, List[dict]]:
    """scenegraph_agent.py:281-295"""
    zones = {name: [] for name, _, _, _ in SPATIAL_ZONES}
    for obj in objects:
        for name, lo, hi, d in SPATIAL_ZONES:
            if obj["direction"] == d and lo <= obj["distance"] < hi:
                zones[name].append(obj)
                break
    return zones


def describe_point_cloud(point_cloud: np.ndarray) -> str:
    """baseline_gpt4o.py:270-287"""
    n = len(point_cloud)
    if n == 0:
        return "LiDAR: No points detected"
    x, y, z = point_cloud[:, 0], point_cloud[:, 1], point_cloud[:, 2]
    return f"""LiDAR Point Cloud Statistics:
- Total points: {n:,}
- X range: [{x.min():.1f}, {x.max():.1f}] m
- Y range: [{y.min():.1f}, {y.max():.1f}] m
- Z range: [{z.min():.1f}, {z.max():.1f}] m
- Average distance from ego: {np.sqrt(x**2 + y**2).mean():.1f} m"""


def describe_annotations(annotations: List[dict]) -> str:
    """baseline_gpt4o.py:289-327"""
    if not annotations:
        return "Annotations: No objects detected"
    categories = Counter(a["category_name"] for a in annotations)
    front = sum(1 for a in annotations if a["translation"][0] > 0)
    left = sum(1 for a in annotations if a["translation"][1] > 0)
    n = len(annotations)
    return f"""Object Annotations:
- Total objects: {n}
- Categories: {dict(categories)}
- Front region: {front} objects
- Back region: {n - front} objects
- Left region: {left} objects
- Right region: {n - left} objects"""


# ----------------------------------------------------------------------------------------------------
# [EXT] devkit-style float64 NumPy versions of the features the reference never implements.
# They are NOT the normative definition (oracle/c/msc_oracle.c is); tests use them as an independent
# cross-check: integer outputs must agree except for points within float rounding of a boundary, float
# outputs to 1e-5.
# ----------------------------------------------------------------------------------------------------
def quat_to_rot(q):
    w, x, y, z = np.asarray(q, np.float64) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def devkit_multisweep(sweeps, min_distance=1.0):
    """App. A.1 from_file_multisweep: sweeps = [(raw (n,5) f32, M 3x4 f64, lag)], returns (4,N) f32 points, (N,) lag."""
    pts, lags = [], []
    for raw, M, lag in sweeps:
        p = raw[:, :4].T.copy()
        keep = ~((np.abs(p[0]) < min_distance) & (np.abs(p[1]) < min_distance))
        p = p[:, keep]
        p[:3] = (np.vstack([M, [0, 0, 0, 1]]) @ np.vstack([p[:3], np.ones(p.shape[1])]))[:3]
        pts.append(p)
        lags.append(np.full(p.shape[1], lag, np.float32))
    return np.hstack(pts), np.concatenate(lags)


def devkit_box_to_frame(box10, poses):
    """App. A.3: translate(-t), rotate(q^-1) for each pose in order; returns (center, R)."""
    c = np.asarray(box10[:3], np.float64).copy()
    R = quat_to_rot(box10[6:10])
    for pose in poses:
        Rp = quat_to_rot(pose[3:])
        c = Rp.T @ (c - np.asarray(pose[:3], np.float64))
        R = Rp.T @ R
    return c, R


def devkit_corners(c, R, wlh):
    w, l, h = wlh
    x = l / 2 * np.array([1, 1, 1, 1, -1, -1, -1, -1.0])
    y = w / 2 * np.array([1, -1, -1, 1, 1, -1, -1, 1.0])
    z = h / 2 * np.array([1, 1, -1, -1, 1, 1, -1, -1.0])
    return R @ np.vstack([x, y, z]) + c[:, None]


def devkit_points_in_box(c, R, wlh, points3n):
    """App. A.2 points_in_box (float64)."""
    corners = devkit_corners(c, R, wlh)
    p1 = corners[:, 0]
    i, j, k = corners[:, 4] - p1, corners[:, 1] - p1, corners[:, 3] - p1
    v = points3n - p1[:, None]
    iv, jv, kv = i @ v, j @ v, k @ v
    return (0 <= iv) & (iv <= i @ i) & (0 <= jv) & (jv <= j @ j) & (0 <= kv) & (kv <= k @ k)


def devkit_box_in_image(c, R, wlh, K, imsize=(1600, 900)):
    """App. A.3 view_points + box_in_image(BoxVisibility.ANY) (float64); returns flag, clipped extent."""
    corners = devkit_corners(c, R, wlh)
    proj = K @ corners
    uv = proj[:2] / proj[2:3]
    vis = (uv[0] > 0) & (uv[0] < imsize[0]) & (uv[1] > 0) & (uv[1] < imsize[1]) & (corners[2] > 1)
    ok = bool(vis.any() and (corners[2] > 0.1).all())
    if not ok:
        return False, np.zeros(4)
    return True, np.array([max(uv[0].min(), 0), max(uv[1].min(), 0), min(uv[0].max(), imsize[0]), min(uv[1].max(), imsize[1])])
