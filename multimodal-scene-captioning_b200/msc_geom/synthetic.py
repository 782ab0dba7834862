"""nuScenes-shaped synthetic samples (SURVEY.md section 8(d) recipe; there is no dataset offline).

`make_sample(idx)` returns the reference loader's sample dict (keys of
/root/reference/src/nuscenes_loader.py:88-101) plus the additive pose / calibration / sweep keys the
multi-sweep path needs.  Seeded with np.random.default_rng(1000 + idx).
"""
from __future__ import annotations

import numpy as np

from .geometry import pose7, quat_to_rot, ref_from_sweep, rot_to_quat, yaw_quat

CAMERA_CHANNELS = ["CAM_FRONT", "CAM_FRONT_RIGHT", "CAM_FRONT_LEFT", "CAM_BACK", "CAM_BACK_LEFT", "CAM_BACK_RIGHT"]
CAMERA_YAWS_DEG = [0.0, -55.0, 55.0, 180.0, 110.0, -110.0]
N_RINGS, N_AZ = 32, 1085  # 34,720 returns per sweep

# (category, w, l, h)
CLASS_TABLE = [
    ("vehicle.car", 1.9, 4.6, 1.7),
    ("vehicle.truck", 2.5, 6.9, 2.8),
    ("vehicle.bus.rigid", 2.9, 11.0, 3.5),
    ("human.pedestrian.adult", 0.67, 0.73, 1.77),
    ("vehicle.bicycle", 0.6, 1.7, 1.3),
    ("movable_object.trafficcone", 0.4, 0.4, 1.07),
    ("movable_object.barrier", 2.5, 0.5, 0.98),
]
VISIBILITY = ["visibility of whole object is between 0 and 40%", "visibility of whole object is between 40 and 60%",
              "visibility of whole object is between 60 and 80%", "visibility of whole object is between 80 and 100%"]

LIDAR_CALIB = pose7([0.94, 0.0, 1.84], yaw_quat(-np.pi / 2))
_R_CAM0 = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])  # camera (x right,y down,z fwd) -> ego


def camera_rig():
    cams = []
    for ch, yaw in zip(CAMERA_CHANNELS, CAMERA_YAWS_DEG):
        a = np.deg2rad(yaw)
        Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
        q = rot_to_quat(Rz @ _R_CAM0)
        t = [1.5 * np.cos(a), 1.5 * np.sin(a) * 0.35, 1.5]
        if ch == "CAM_BACK":
            K = [[809.2, 0, 829.2], [0, 809.2, 481.8], [0, 0, 1.0]]
        else:
            K = [[1266.4, 0, 816.3], [0, 1266.4, 491.5], [0, 0, 1.0]]
        cams.append({"channel": ch, "calib": pose7(t, q), "intrinsic": np.array(K, dtype=np.float64)})
    return cams


def _sweep_ranges(rng, n):
    ring = np.tile(np.arange(N_RINGS, dtype=np.int64), N_AZ)[:n]
    az_idx = np.repeat(np.arange(N_AZ, dtype=np.int64), N_RINGS)[:n]
    elev = np.deg2rad(-30.67 + (41.34 / (N_RINGS - 1)) * ring)
    az = (2 * np.pi / N_AZ) * az_idx
    free = rng.uniform(2.0, 80.0, n)
    with np.errstate(divide="ignore"):
        ground = np.where(elev < 0, 1.84 / np.tan(-elev) / np.cos(elev), np.inf)
    rang = np.minimum(ground, free) + rng.normal(0.0, 0.02, n)
    x = rang * np.cos(elev) * np.sin(az)  # sensor frame: x right, y forward, z up
    y = rang * np.cos(elev) * np.cos(az)
    z = rang * np.sin(elev)
    return np.stack([x, y, z], 1), ring


def make_sample(idx: int, n_sweeps: int = 10, n_boxes=None, box_point_fraction: float = 0.25,
                integer_intensity: bool = False, with_images: bool = False) -> dict:
    rng = np.random.default_rng(1000 + idx)
    if n_boxes is None:
        n_boxes = 60
    elif n_boxes == "mini":
        n_boxes = int(rng.integers(60, 121))
    # ---- ego trajectory (keyframe = sweep 0, earlier sweeps go back in time) ----
    origin = rng.uniform(300.0, 2000.0, 2)
    speed = rng.uniform(0.0, 15.0)
    yaw0 = rng.uniform(-np.pi, np.pi)
    drift = rng.uniform(-0.02, 0.02)
    dt = 0.05
    ego_poses = []
    pos = np.array([origin[0], origin[1], 0.0])
    yaw = yaw0
    for s in range(n_sweeps):
        ego_poses.append(pose7(pos.copy(), yaw_quat(yaw)))
        pos = pos - speed * dt * np.array([np.cos(yaw), np.sin(yaw), 0.0])
        yaw = yaw - drift
    ref_ego = ego_poses[0]
    # ---- boxes in the keyframe ego frame -> global ----
    Rg = quat_to_rot(ref_ego[3:])
    cls = rng.integers(0, len(CLASS_TABLE), n_boxes)
    annotations, boxes_lidar = [], []
    Rl = quat_to_rot(LIDAR_CALIB[3:])
    for b in range(n_boxes):
        name, w, l, h = CLASS_TABLE[cls[b]]
        scale = rng.uniform(0.9, 1.1)
        c_ego = np.array([rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(-1.5, 0.5)])
        byaw = rng.uniform(-np.pi, np.pi)
        Rb = quat_to_rot(yaw_quat(byaw))
        c_glob = Rg @ c_ego + ref_ego[:3]
        q_glob = rot_to_quat(Rg @ Rb)
        v = rng.uniform(-6.0, 6.0, 2) * (rng.uniform() < 0.6)
        annotations.append({
            "token": f"synth_ann_{idx}_{b}", "category_name": name, "instance_token": f"synth_inst_{idx}_{b}",
            "translation": [float(c_glob[0]), float(c_glob[1]), float(c_glob[2])],
            "size": [float(w * scale), float(l * scale), float(h * scale)],
            "rotation": [float(q) for q in q_glob], "velocity": [float(v[0]), float(v[1])],
            "attribute_tokens": ["vehicle.moving"] if name.startswith("vehicle") else [],
            "visibility_token": VISIBILITY[int(rng.integers(0, 4))], "num_lidar_pts": 0, "num_radar_pts": 0,
        })
        # same box in the keyframe lidar frame, for placing returns inside it
        c_l = Rl.T @ (c_ego - LIDAR_CALIB[:3])
        boxes_lidar.append((c_l, Rl.T @ Rb, np.array([l, w, h]) * scale))
    # ---- sweeps ----
    sweeps = []
    n = N_RINGS * N_AZ
    for s in range(n_sweeps):
        M = ref_from_sweep(ref_ego, LIDAR_CALIB, ego_poses[s], LIDAR_CALIB)
        xyz, ring = _sweep_ranges(rng, n)
        if n_boxes > 0 and box_point_fraction > 0:
            sel = np.nonzero(rng.uniform(size=n) < box_point_fraction)[0]
            which = rng.integers(0, n_boxes, sel.size)
            local = rng.uniform(-0.5, 0.5, (sel.size, 3))
            C = np.stack([boxes_lidar[w_][0] for w_ in range(n_boxes)])
            Rm = np.stack([boxes_lidar[w_][1] for w_ in range(n_boxes)])
            D = np.stack([boxes_lidar[w_][2] for w_ in range(n_boxes)])
            p_ref = C[which] + np.einsum("nij,nj->ni", Rm[which], local * D[which])
            Minv_R = M[:, :3].T
            xyz[sel] = (p_ref - M[:, 3]) @ Minv_R.T  # back into sweep s's own sensor frame
        inten = rng.uniform(0.0, 255.0, n)
        if integer_intensity:
            inten = np.floor(inten)
        raw = np.empty((n, 5), dtype=np.float32)
        raw[:, :3] = xyz
        raw[:, 3] = inten
        raw[:, 4] = ring
        sweeps.append({"points_raw": raw, "ref_from_sensor": M, "ego_pose": ego_poses[s], "calib": LIDAR_CALIB,
                       "time_lag": float(s * dt)})
    cams = camera_rig()
    for c in cams:
        c["ego_pose"] = ref_ego  # synthetic rig: cameras share the keyframe ego pose
    sample = {
        "sample_token": f"synth_sample_{idx:06d}", "timestamp": 1532402927647951 + idx * 500000,
        "scene_description": "Synthetic nuScenes-shaped scene", "scene_name": f"scene-s{idx // 40:04d}",
        "images": [], "camera_names": list(CAMERA_CHANNELS),
        "point_cloud": sweeps[0]["points_raw"][:, :4],  # (N,4) view, 20-byte pitch, like the devkit loader
        "annotations": annotations, "metadata": {"location": "synthetic", "nbr_objects": len(annotations)},
        # additive keys (new pose/calibration data as extra keys only)
        "lidar_sweeps": sweeps, "ego_pose": ref_ego, "lidar_calib": LIDAR_CALIB, "cameras": cams,
    }
    if with_images:
        sample["images"] = [rng.integers(0, 255, (900, 1600, 3), dtype=np.uint8) for _ in range(6)]
    return sample


def edge_case_cloud() -> np.ndarray:
    """Keyframe cloud with points exactly on every threshold the reference tests with strict compares
    (lidar_agent.py:106-110, :128) and on BEV cell edges (:548-551)."""
    rows = []
    for d in [1.0, np.nextafter(np.float32(1.0), np.float32(2.0)), 50.0, np.nextafter(np.float32(50.0), np.float32(0.0)),
              49.999996, 0.99999994, 25.0]:
        for ang in [0.0, 45.0, 90.0, 135.0, 180.0, 225.0, 270.0, 315.0, 22.5, 337.5]:
            a = np.deg2rad(ang)
            for z in [-3.0, -2.9999998, 5.0, 4.9999995, -1.4, -1.4000001, -1.3999999, 0.0, -0.0, 1.0]:
                rows.append([d * np.cos(a), d * np.sin(a), z, 7.0])
    for c in np.arange(-50.0, 50.01, 0.125):  # 800-grid cell edges
        rows.append([c, 3.0, 0.5, 1.0])
        rows.append([3.0, c, -2.0, 2.0])
        rows.append([np.nextafter(np.float32(c), np.float32(100.0)), -7.0, 2.0, 3.0])
        rows.append([-7.0, np.nextafter(np.float32(c), np.float32(-100.0)), -1.5, 4.0])
    return np.asarray(rows, dtype=np.float32)
