"""SURVEY.md section 8(d) CPU baseline: the reference's OWN functions + the NumPy [EXT] restatement, on this host's cores.

Test / measurement infrastructure (lives under oracle/, never imported by the product).  Per sample of BASELINE config 3
(10 sweeps x 34,720 points, 60 boxes, 6 cameras):
    [EXT] devkit-style multi-sweep aggregation (oracle/numpy_ref.py devkit_multisweep, float64 NumPy)
    reference  LiDARAgent._preprocess_point_cloud + _segment_ground + _generate_multi_layer_bev   (lidar_agent.py:103-132, 532-642, verbatim:
               imported from /root/reference/src, Python per-point loops and all)
    [EXT] 200x200 BEV (count / height / intensity, NumPy ufunc.at), points-in-box count / nearest / centroid for every box,
          box -> 6-camera projection (devkit_* functions)
DBSCAN and every LLM call are excluded.  multiprocessing.Pool(P) over samples, BLAS threads pinned to 1,
perf_counter over >= 32 samples after 2 warm-ups, for P = os.cpu_count() and P = 1.  Needs /root/reference (build container only); the
result line is committed under profiles/ and quoted in BASELINE.md.

    PYTHONDONTWRITEBYTECODE=1 python oracle/cpu_baseline_reference.py [--samples 32] [--out profiles/r2_cpu_baseline_reference_python.json]
"""
import argparse
import json
import os
import sys
import time

for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[v] = "1"
sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.join(ROOT, "multimodal-scene-captioning_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

_AGENT = None


def _agent():
    global _AGENT
    if _AGENT is None:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            from agents.content_transform.lidar_agent import LiDARAgent
        _AGENT = LiDARAgent(object(), "m", "n")  # geometry methods never touch the client (src/export_sample_data.py:53-65)
    return _AGENT


def one_sample(idx: int) -> int:
    from msc_geom.synthetic import make_sample
    from oracle import numpy_ref as R
    s = make_sample(idx)
    t0 = time.perf_counter()
    agent = _agent()
    pts, _ = R.devkit_multisweep([(sw["points_raw"], sw["ref_from_sensor"], sw["time_lag"]) for sw in s["lidar_sweeps"]])
    cloud = np.ascontiguousarray(pts.T.astype(np.float32))
    kept = agent._preprocess_point_cloud(cloud)                         # reference, verbatim
    ground, obj = agent._segment_ground(kept)                           # reference, verbatim
    agent._generate_multi_layer_bev(ground, obj)                        # reference, verbatim (800x800, Python per-point loops)
    x, y, z, inten = kept[:, 0], kept[:, 1], kept[:, 2], kept[:, 3]
    ix, iy = R.to_pixels(kept[:, :2], 50, 200)                          # [EXT] 200x200 grid
    cnt = np.zeros((200, 200), np.int64); np.add.at(cnt, (iy, ix), 1)
    hgt = np.zeros((200, 200), np.float32); np.maximum.at(hgt, (iy, ix), z)
    isum = np.zeros((200, 200)); np.add.at(isum, (iy, ix), inten.astype(np.float64))
    P3 = np.vstack([x, y, z]).astype(np.float64)
    members = 0
    for ann in s["annotations"]:                                        # [EXT] membership + projection
        box = np.array(ann["translation"] + ann["size"] + ann["rotation"])
        c, Rm = R.devkit_box_to_frame(box, [s["ego_pose"], s["lidar_calib"]])
        m = R.devkit_points_in_box(c, Rm, ann["size"], P3)
        if m.any():
            members += int(m.sum())
            float(np.sqrt(x[m] ** 2 + y[m] ** 2).min()); P3[:, m].mean(1)
        for cam in s["cameras"]:
            c2, R2 = R.devkit_box_to_frame(box, [cam["ego_pose"], cam["calib"]])
            R.devkit_box_in_image(c2, R2, ann["size"], cam["intrinsic"])
    return time.perf_counter() - t0


def run(n_samples: int, procs: int) -> dict:
    import multiprocessing as mp
    idx = [1000 + (i % 8) for i in range(n_samples)]
    if procs == 1:
        for i in idx[:2]:
            one_sample(i)
        t0 = time.perf_counter()
        inner = [one_sample(i) for i in idx]
        dt = time.perf_counter() - t0
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            pool.map(one_sample, idx[:2 * procs])                       # warm-ups in every worker
            t0 = time.perf_counter()
            inner = pool.map(one_sample, idx, chunksize=1)
            dt = time.perf_counter() - t0
    # samples_per_s: wall clock of the pool (includes generating the synthetic sample in the worker, ~10 %); samples_per_s_compute:
    # processes / mean time inside the timed functions only
    return {"processes": procs, "samples": n_samples, "wall_s": round(dt, 3), "samples_per_s": n_samples / dt,
            "mean_s_per_sample_in_worker": float(np.mean(inner)), "samples_per_s_compute": procs / float(np.mean(inner))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=32)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_cpu_baseline_reference_python.json"))
    a = ap.parse_args()
    cores = os.cpu_count() or 1
    one_sample(1000)  # generator / import warm-up in the parent (forked workers inherit it)
    res = {"what": "reference LiDARAgent functions verbatim (a3, a4, a8) + NumPy [EXT] restatement, BASELINE config 3, DBSCAN and LLM excluded",
           "host": {"cpu_count": cores, "where": "build container (the reference tree does not exist on the GPU box)"},
           "pool": run(a.samples, cores), "single": run(max(8, a.samples // 4), 1), "unit": "samples/s"}
    res["value"] = res["pool"]["samples_per_s_compute"]
    res["value_1_process"] = res["single"]["samples_per_s_compute"]
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
