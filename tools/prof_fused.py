"""Small driver for ncu: a few launches of the fused kernel on a 148-sample config-3 batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-scene-captioning_b200")); sys.path.insert(0, ROOT)
import torch
from msc_geom.synthetic import make_sample
from msc_geom.layout import pack_batch, tile_batch, GeomParams
from msc_geom.engine import GeometryEngine
from msc_geom import _capi
n_unique = int(os.environ.get("PROF_UNIQUE", "4")); reps = int(os.environ.get("PROF_REPS", "37")); launches = int(os.environ.get("PROF_LAUNCHES", "3"))
eng = GeometryEngine()
_capi.set_option("fov", int(os.environ.get("PROF_FOV", "1")))
_capi.set_option("grid", int(os.environ.get("PROF_GRID", "0")))
_capi.set_option("config", int(os.environ.get("PROF_CONFIG", "10")))
_capi.set_option("ppt", int(os.environ.get("PROF_PPT", "2")))
hb = tile_batch(pack_batch([make_sample(i) for i in range(n_unique)]), reps)
db = eng.upload(hb); out = eng.alloc_result(hb)
for _ in range(launches):
    eng.run_fused(db, out)
torch.cuda.synchronize()
print("ok", hb.n_samples, "samples")
