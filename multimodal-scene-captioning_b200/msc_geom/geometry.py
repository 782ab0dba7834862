"""Host-side float64 rigid-transform helpers (input preparation, not the hot path).

Restates the pose algebra of nuscenes-devkit (un-vendored dependency of the reference,
/root/reference/requirements.txt:4) that `LidarPointCloud.from_file_multisweep` uses
(SURVEY.md Appendix A.1).  Quaternions are [w, x, y, z] as in the loader's annotation dicts
(/root/reference/src/nuscenes_loader.py:185).
"""
from __future__ import annotations

import functools

import numpy as np


def quat_to_rot(q) -> np.ndarray:
    w, x, y, z = (float(v) for v in q)
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ],
        dtype=np.float64,
    )


def rot_to_quat(R) -> np.ndarray:
    """Rotation matrix -> unit quaternion [w,x,y,z] (Shepperd's method)."""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    q = np.array(q, dtype=np.float64)
    return q / np.linalg.norm(q)


def yaw_quat(yaw: float) -> np.ndarray:
    return np.array([np.cos(yaw / 2), 0.0, 0.0, np.sin(yaw / 2)], dtype=np.float64)


def transform_matrix(translation, rotation, inverse: bool = False) -> np.ndarray:
    """devkit `transform_matrix(t, Quaternion(q), inverse)`: [[R, t],[0,1]] or its inverse."""
    R = quat_to_rot(rotation)
    t = np.asarray(translation, dtype=np.float64)
    T = np.eye(4)
    if inverse:
        T[:3, :3] = R.T
        T[:3, 3] = R.T @ (-t)
    else:
        T[:3, :3] = R
        T[:3, 3] = t
    return T


def pose7(translation, rotation) -> np.ndarray:
    return np.concatenate([np.asarray(translation, np.float64), np.asarray(rotation, np.float64)])


def ref_from_sweep(ref_ego_pose7, ref_calib7, sweep_ego_pose7, sweep_calib7) -> np.ndarray:
    """App. A.1: M = ref_from_car @ car_from_global @ global_from_car(s) @ car_from_current(s); 3x4 f64."""
    ref_from_car = transform_matrix(ref_calib7[:3], ref_calib7[3:], inverse=True)
    car_from_global = transform_matrix(ref_ego_pose7[:3], ref_ego_pose7[3:], inverse=True)
    global_from_car = transform_matrix(sweep_ego_pose7[:3], sweep_ego_pose7[3:], inverse=False)
    car_from_current = transform_matrix(sweep_calib7[:3], sweep_calib7[3:], inverse=False)
    M = ref_from_car @ car_from_global @ global_from_car @ car_from_current
    return np.ascontiguousarray(M[:3, :4])


@functools.lru_cache(maxsize=64)
def sqrt_thresholds(lo: float, hi: float):
    """Smallest float32 s with sqrt(s) > lo and largest float32 s with sqrt(s) < hi (float32 sqrt).

    The kernels compare s = x*x + y*y against these instead of taking a square root per point; float32
    sqrt is correctly rounded and monotonic, so the predicate is identical to the reference's
    `distances > lo` / `distances < hi` (lidar_agent.py:106-107)."""
    lo32, hi32 = np.float32(lo), np.float32(hi)

    def first_true(pred):
        """Smallest non-negative float32 (by bit pattern, which orders non-negative floats) satisfying a monotone predicate."""
        a, b = 0, 0x7F800000  # +0.0 .. +inf
        if not pred(np.uint32(b).view(np.float32)):
            return None
        while a < b:
            m = (a + b) // 2
            if pred(np.uint32(m).view(np.float32)):
                b = m
            else:
                a = m + 1
        return np.uint32(a).view(np.float32)

    with np.errstate(invalid="ignore"):
        s_lo = first_true(lambda s: np.sqrt(s) > lo32)          # monotone: false ... true
        above = first_true(lambda s: not (np.sqrt(s) < hi32))   # first s that is NOT below the upper limit
    s_lo = np.float32(np.inf) if s_lo is None else s_lo
    if above is None:
        s_hi = np.float32(np.inf)
    elif above.view(np.uint32) == 0:
        s_hi = np.float32(-1.0)  # nothing passes
    else:
        s_hi = np.uint32(above.view(np.uint32) - 1).view(np.float32)
    return float(s_lo), float(s_hi)
