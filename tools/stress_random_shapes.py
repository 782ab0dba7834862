"""One-off stress of the streaming kernel against the oracle: the loop of tests/test_gpu_parity.py::test_fused_stream4_random_shapes_and_options
with other seeds, more sweeps per sample (beyond the staged poses), 0 / 6 / 8 cameras and more boxes.  usage: stress_random_shapes.py <seed> [<seed> ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-scene-captioning_b200")); sys.path.insert(0, ROOT)
import numpy as np
from msc_geom import _capi
from msc_geom.engine import GeometryEngine
from msc_geom.layout import GeomParams
from msc_geom.synthetic import make_sample
from tests.test_gpu_parity import check_fused

eng = GeometryEngine()
t0 = time.time(); n_ok = 0
for seed in map(int, sys.argv[1:] or ["1"]):
    rng = np.random.default_rng(seed)
    for trial in range(8):
        n = int(rng.integers(1, 5))
        s = [make_sample(int(rng.integers(500, 900)), n_sweeps=int(rng.integers(1, 14)), n_boxes=int(rng.integers(0, 130))) for _ in range(n)]
        for smp in s:
            for sw in smp["lidar_sweeps"]:
                sw["points_raw"] = sw["points_raw"][: int(rng.integers(0, 34720))]
        params = None
        if trial % 3 == 2:
            params = GeomParams(range_max=float(rng.choice([30.0, 40.0, 50.0])), bev_range=float(rng.choice([32.0, 51.2, 60.0])),
                                bev_res=int(rng.choice([64, 128, 200, 256])), z_max=3.0, ground_z=-1.2)
        ppt, grid, window, cull_shift = int(rng.choice([2, 4])), int(rng.choice([0, 1, 2, 7, 29, 148])), int(rng.choice([0, 0, 16, 40, 60])), int(rng.choice([-1, -1, 1, 3]))
        if cull_shift == 1 and (ppt == 2 or (params is not None and params.bev_res > 200)):
            cull_shift = 2
        opts = {"ppt": ppt, "grid": grid, "window": window, "cull_shift": cull_shift, "standard": int(rng.integers(0, 2))}
        for k, v in opts.items():
            _capi.set_option(k, v)
        n_cams = int(rng.choice([0, 6, 6, 8]))
        import dataclasses
        params = dataclasses.replace(params or GeomParams(), n_cams=n_cams)
        try:
            check_fused(eng, s, params=params, config=10, n_cams=n_cams)
            n_ok += 1
        except Exception as e:  # noqa: BLE001 -- report and carry on: this is a survey, not a gate
            print("FAIL seed", seed, "trial", trial, opts, "cams", n_cams, [(len(x["lidar_sweeps"]), len(x["annotations"])) for x in s], repr(e)[:300], flush=True)
print("trials ok:", n_ok, "seconds:", round(time.time() - t0, 1))
