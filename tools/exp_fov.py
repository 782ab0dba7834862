"""Timing experiment: the streaming kernels with and without the FOV counts (option fov) on the bench batch."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-scene-captioning_b200")); sys.path.insert(0, ROOT)
import torch
from msc_geom.synthetic import make_sample
from msc_geom.layout import pack_batch, tile_batch
from msc_geom.engine import GeometryEngine
from msc_geom import _capi
eng = GeometryEngine()
hb = tile_batch(pack_batch([make_sample(i) for i in range(8)]), 74)
db = eng.upload(hb); out = eng.alloc_result(hb)
for cfg in (7, 9):
    for fov in (1, 0):
        for win in (0, 64):
            _capi.set_option("config", cfg); _capi.set_option("fov", fov); _capi.set_option("window", win)
            for _ in range(3): eng.run_fused(db, out)
            torch.cuda.synchronize()
            _capi.set_option("time_kernel", 1)
            for _ in range(10): eng.run_fused(db, out)
            torch.cuda.synchronize()
            kt = _capi.kernel_times(10); _capi.set_option("time_kernel", 0)
            print(json.dumps({"config": cfg, "fov": fov, "window": _capi.get_option("last_window"), "kernel_ms": round(sum(kt) / len(kt), 4)}), flush=True)
