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


# ----------------------------------------------------------------------------------------------------
# cluster 4-view raster and batch mosaic
# ----------------------------------------------------------------------------------------------------
_CIRCLE_R2 = [(dx, dy) for dy in range(-2, 3) for dx in range(-(2 - abs(dy)), 2 - abs(dy) + 1)]  # cv2.circle(r=2, filled)


def cluster_view_params(points: np.ndarray, img_size: int = 256):
    """lidar_agent.py:255-264: cluster mean (float32, NumPy's own reduction) and the pixel scale."""
    center = points[:, :3].mean(axis=0)
    centered = points[:, :3] - center
    max_range = max(centered[:, 0].max() - centered[:, 0].min(), centered[:, 1].max() - centered[:, 1].min(),
                    centered[:, 2].max() - centered[:, 2].min())
    scale = (img_size * 0.35) / max_range if max_range > 0 else 1
    return center, centered, scale


def cluster_view_points(points: np.ndarray, img_size: int = 256) -> np.ndarray:
    """Point discs of _generate_cluster_visualization (lidar_agent.py:267-351) without cv2: white 2x2 grid, per view the
    points in array order, each a filled radius-2 circle in the view-normalised intensity grey."""
    _, centered, scale = cluster_view_params(points, img_size)
    grid = np.ones((img_size * 2, img_size * 2, 3), dtype=np.uint8) * 255

    def splat(px, py, valid, qx, qy):
        px, py = px[valid], py[valid]
        inten = points[valid, 3]
        inten = ((inten - inten.min()) / (inten.max() - inten.min() + 1e-6) * 255).astype(np.uint8)
        for x, y, g in zip(px, py, inten):
            cx, cy = qx * img_size + x, qy * img_size + img_size - y - 1
            for dx, dy in _CIRCLE_R2:
                xx, yy = cx + dx, cy + dy
                if 0 <= xx < 2 * img_size and 0 <= yy < 2 * img_size:
                    grid[yy, xx] = g

    for (a1, a2, qx, qy) in ((0, 1, 0, 0), (0, 2, 1, 0), (1, 2, 0, 1)):
        px = (centered[:, a1] * scale + img_size / 2).astype(int)
        py = (centered[:, a2] * scale + img_size / 2).astype(int)
        splat(px, py, (px >= 0) & (px < img_size) & (py >= 0) & (py < img_size), qx, qy)
    ang = np.pi / 6
    rot_x = np.array([[1, 0, 0], [0, np.cos(ang), -np.sin(ang)], [0, np.sin(ang), np.cos(ang)]])
    rot_y = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    rot = centered @ rot_x.T @ rot_y.T
    ix = ((rot[:, 0] + rot[:, 1] * 0.5) * scale + img_size / 2).astype(int)
    iy = ((rot[:, 2] - rot[:, 1] * 0.5) * scale + img_size / 2).astype(int)
    splat(ix, iy, (ix >= 0) & (ix < img_size) & (iy >= 0) & (iy < img_size), 1, 1)
    return grid


def cluster_view_overlays(grid: np.ndarray, img_size: int = 256) -> np.ndarray:
    """Axes and titles (lidar_agent.py:299-313, :353-354).  Drawn after all discs: a disc bleeds at most 2 px over a
    quadrant border, the overlays sit >= 10 px inside their quadrant, so the reference's interleaved order gives the same image."""
    import cv2
    grid = np.ascontiguousarray(grid)
    for (qx, qy, title) in ((0, 0, "Top (XY)"), (1, 0, "Side (XZ)"), (0, 1, "Front (YZ)")):
        ox, oy, c = qx * img_size, qy * img_size, img_size // 2
        cv2.line(grid, (ox + c, oy + c), (ox + c + 30, oy + c), (0, 0, 255), 2)
        cv2.line(grid, (ox + c, oy + c), (ox + c, oy + c - 30), (0, 255, 0), 2)
        cv2.putText(grid, title, (ox + 10, oy + 20), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (0, 0, 0), 1)
    cv2.putText(grid, "3D View", (img_size + 10, img_size + 20), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (0, 0, 0), 1)
    return grid


def generate_cluster_visualization(points: np.ndarray, img_size: int = 256) -> np.ndarray:
    """lidar_agent.py:241-356"""
    return cluster_view_overlays(cluster_view_points(points, img_size), img_size)


def cluster_mosaic(images) -> np.ndarray:
    """lidar_agent.py:366-386: up to three columns, '#idx' labels; a single image is passed through."""
    import cv2
    if len(images) == 1:
        return images[0]
    n = len(images)
    cols = min(3, n)
    rows = (n + cols - 1) // cols
    h, w = images[0].shape[:2]
    out = np.ones((rows * h, cols * w, 3), dtype=np.uint8) * 255
    for idx, img in enumerate(images):
        r, c = idx // cols, idx % cols
        out[r * h:(r + 1) * h, c * w:(c + 1) * w] = img
        cv2.putText(out, f"#{idx}", (c * w + 10, r * h + 50), cv2.FONT_HERSHEY_SIMPLEX, 1.5, (255, 0, 0), 3)
    return out
