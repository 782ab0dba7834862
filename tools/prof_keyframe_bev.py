import os, sys, time
ROOT='/root/repo'
sys.path.insert(0, os.path.join(ROOT, "multimodal-scene-captioning_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from msc_geom import ops
from msc_geom.engine import GeometryEngine
from msc_geom.layout import GeomParams
from msc_geom.lidar_agent import finish_bev_layers
from msc_geom.synthetic import make_sample
eng = GeometryEngine()
def med(f, n=9):
    for _ in range(3): f()
    ts=[]
    for _ in range(n):
        torch.cuda.synchronize(); t=time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter()-t)
    return sorted(ts)[n//2]*1e3
for name, nsw in (("K1",1),("K10",10)):
    s = make_sample(3, n_sweeps=nsw)
    sweeps = [(sw["points_raw"], sw["ref_from_sensor"], sw["time_lag"]) for sw in s["lidar_sweeps"]]
    xyzi,_ = ops.aggregate_sweeps(eng, sweeps)
    kept, ground, obj = ops.keyframe_filter_split(eng, xyzi, GeomParams(bev_res=800))
    layers = ops.keyframe_bev_layers(eng, ground, obj, 800, 50.0)
    t_gpu = med(lambda: ops.keyframe_bev_layers(eng, ground, obj, 800, 50.0))
    t_host = med(lambda: finish_bev_layers(*[a.copy() for a in layers], 800, 50.0))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    print(name, len(ground), len(obj), "raster+copies ms %.3f" % t_gpu, "host finish ms %.3f" % t_host, "count max", int(layers[0].max()))
