"""Regenerates the golden fixtures in this directory by IMPORTING THE REFERENCE ITSELF
(/root/reference/src, read-only) and running its own functions on seeded inputs.  Run in the build
container only (the reference does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The Azure client is never touched: geometry methods do not use self.client (the reference's own export scripts
construct LiDARAgent(MockClient(), ...) the same way, src/export_sample_data.py:53-65) and call_llm is stubbed
where process() needs it.
"""
import csv
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.join(ROOT, "multimodal-scene-captioning_b200"))

from agents.base_agent import BaseAgent  # noqa: E402
from agents.content_transform.lidar_agent import LiDARAgent  # noqa: E402
from agents.content_transform.scenegraph_agent import SceneGraphAgent  # noqa: E402
from baseline_gpt4o import RawGPT4oBaseline  # noqa: E402
from nuscenes_loader import create_loader  # noqa: E402

from msc_geom.synthetic import edge_case_cloud, make_sample  # noqa: E402


def lidar_case(agent, pc, name):
    kept = agent._preprocess_point_cloud(pc)
    ground, obj = agent._segment_ground(kept)
    bev = agent._generate_multi_layer_bev(ground, obj)
    np.savez_compressed(os.path.join(HERE, f"keyframe_{name}.npz"), points=np.ascontiguousarray(pc), kept=kept, ground=ground, object=obj,
                        semantic=bev["semantic"], height=bev["height"], density=bev["density"])
    print(name, pc.shape, kept.shape, ground.shape, obj.shape)
    return kept, ground, obj


def main():
    agent = LiDARAgent(object(), "m", "n")
    # (1) the mock loader's sample, seeded like BASELINE config 1
    np.random.seed(0)
    sample = create_loader(None, use_mock=True).get_sample_by_scene_index(0, 0)
    lidar_case(agent, sample["point_cloud"], "mock")
    # (2) points exactly on every threshold / cell edge
    lidar_case(agent, edge_case_cloud(), "edge")
    # (3) a nuScenes-shaped synthetic keyframe handed over as the devkit's 20-byte-pitch view (subsampled 1:4)
    raw = make_sample(0, n_sweeps=1)["lidar_sweeps"][0]["points_raw"][::4].copy()
    _, _, obj = lidar_case(agent, raw[:, :4], "synth")
    # (4) degenerate clouds
    lidar_case(agent, np.zeros((0, 4), np.float32), "empty")
    flat = np.array([[3, 4, -2.0, 1], [5, 5, -1.0, 2], [6, -7, -1.0, 3], [-8, 2, -1.0, 4]], np.float32)  # all object heights equal
    lidar_case(agent, flat, "flatobj")

    # (5) cluster metadata: run the reference's own DBSCAN + metadata code with the LLM/visualisation stubbed
    captured = {}
    agent._generate_cluster_visualization = lambda pts: np.zeros((8, 8, 3), np.uint8)
    def fake_classify(vis, meta):
        captured.setdefault("meta", []).extend(meta)
        return [{"category": "car", "confidence": 0.9} for _ in meta]
    agent._classify_batch_with_llm = fake_classify
    from sklearn.cluster import DBSCAN
    labels = DBSCAN(eps=agent.dbscan_eps, min_samples=agent.dbscan_min_samples).fit(obj[:, :3]).labels_
    objs = agent._detect_objects_3d(obj)
    meta = captured.get("meta", [])
    np.savez_compressed(os.path.join(HERE, "clusters_synth.npz"), object=obj, labels=labels.astype(np.int32),
                        center=np.array([m["center"] for m in meta], np.float32).reshape(-1, 3),
                        dimensions=np.array([m["dimensions"] for m in meta], np.float32).reshape(-1, 3),
                        distance=np.array([m["distance"] for m in meta], np.float32),
                        num_points=np.array([m["num_points"] for m in meta], np.int64),
                        direction=np.array([m["direction"] for m in meta]))
    print("clusters", len(meta), "detected", len(objs))
    # (5b) cluster 4-view rasters and the batch mosaic, straight from the reference
    ref_agent = LiDARAgent(object(), "m", "n")
    rng = np.random.default_rng(7)
    view_clusters = [obj[labels == l] for l in sorted(set(labels.tolist()) - {-1})[:3]]
    car = rng.uniform(-0.5, 0.5, (900, 3)) * np.array([4.6, 1.9, 1.7]) + np.array([12.0, -3.0, -1.0])
    view_clusters.append(np.concatenate([car, rng.integers(0, 255, (900, 1))], 1).astype(np.float32))     # dense box-shaped cluster
    view_clusters.append(np.tile(np.array([[5.0, 5.0, 0.5, 10.0]], np.float32), (6, 1)))                  # all points identical: max_range == 0
    line = np.zeros((40, 4), np.float32); line[:, 0] = np.linspace(-3, 3, 40); line[:, 3] = 77.0         # constant intensity, thin line
    view_clusters.append(line)
    views = [ref_agent._generate_cluster_visualization(c) for c in view_clusters]
    captured_img = {}
    import agents.content_transform.lidar_agent as _la_mod
    ref_agent._image_to_base64 = lambda img: captured_img.setdefault("mosaic", img.copy()) is None or "b64"
    ref_agent.call_llm = lambda *a, **k: "{}"
    try:
        ref_agent._classify_batch_with_llm(views[:5], [{"index": i, "center": np.zeros(3), "dimensions": np.ones(3), "distance": 1.0,
                                                        "direction": "front", "num_points": 5} for i in range(5)])
    except Exception as e:  # the stubbed reply is not a classification; the mosaic is captured before it is parsed
        print("classify stub:", type(e).__name__)
    np.savez_compressed(os.path.join(HERE, "cluster_views.npz"), n=len(view_clusters), mosaic=captured_img["mosaic"],
                        **{f"pts_{i}": c for i, c in enumerate(view_clusters)}, **{f"img_{i}": v for i, v in enumerate(views)})
    print("cluster views", [v.shape for v in views], captured_img["mosaic"].shape)
    angles = np.concatenate([np.arange(0, 360, 7.5), [22.5, 67.5, 112.5, 157.5, 202.5, 247.5, 292.5, 337.5, 359.999]])
    dirs = [agent._get_direction(np.array([np.cos(np.deg2rad(a)), np.sin(np.deg2rad(a))], np.float32) * np.float32(12.5)) for a in angles]

    # (6) full LiDARAgent.process on the mock sample with call_llm stubbed
    BaseAgent.call_llm = lambda self, messages, temperature=0.7, max_retries=8, response_format=None: "STUB"
    agent2 = LiDARAgent(object(), "m", "LiDARAgent")
    agent2._generate_cluster_visualization = lambda pts: np.zeros((8, 8, 3), np.uint8)
    agent2._classify_batch_with_llm = lambda vis, meta: [{"category": "car", "confidence": 0.9} for _ in meta]
    out = agent2.process(sample["point_cloud"])
    sf = dict(out["semantic_features"])
    near = sf.pop("nearest_object")
    sf["nearest_object_distance"] = None if near is None else float(near.distance)
    process_golden = {"bev_metadata": out["bev_metadata"], "semantic_features": sf, "structured_report": out["structured_report"],
                      "detected_objects": out["detected_objects"], "observations": out["observations"]}

    # (7) annotation path: mock annotations, the three shipped annotations.csv tables, boundary cases
    sg = SceneGraphAgent(object(), "m", "SceneGraphAgent")
    cases = {"mock": sample["annotations"]}
    for i, d in enumerate(["scene_1_ca9a282c_assets", "scene_2_3e8750f3_assets", "scene_3_8687ba92_assets"], start=1):
        rows = list(csv.DictReader(open(f"/root/reference/docs/assets/{d}/annotations.csv")))
        cases[f"docs_scene_{i}"] = [{"category_name": r["category"], "translation": [float(r["x"]), float(r["y"]), float(r["z"])],
                                     "size": [float(r["width"]), float(r["length"]), float(r["height"])], "rotation": [1.0, 0.0, 0.0, 0.0],
                                     "velocity": [0.0, 0.0], "attribute_tokens": [], "visibility_token": r["visibility"]} for r in rows]
    edge = []
    for m in [1.0, 10.0, 7.25, 353.794, 29.999999999999996, 30.0, 9.999999999999998, 50.0]:
        for sx, sy in [(1, 1), (-1, 1), (-1, -1), (1, -1), (1, 0), (0, 1), (-1, 0), (0, -1)]:
            edge.append({"category_name": "vehicle.car", "translation": [sx * m, sy * m, 0.0], "velocity": [0.3, 0.4]})
    edge += [{"category_name": "human.pedestrian.child", "translation": [0.0, 0.0, 0.0], "velocity": [float("nan"), 1.0]},
             {"category_name": "movable_object.trafficcone", "translation": [3.0, 4.0, 0.0], "velocity": None},
             {"category_name": "static_object.bicycle_rack", "translation": [-0.0, 5.0, 0.0], "velocity": [0.5, 0.0]},
             {"category_name": "vehicle.construction", "translation": [6.0, -8.0, 0.0], "velocity": [0.30000000000000004, 0.4]},
             {"category_name": "movable_object.barrier", "translation": [-30.0, 1e-9, 0.0], "velocity": [None, 1.0], "visibility_token": "v40-60"},
             {"translation": [7.0710678118654755, 7.0710678118654755, 0.0]}]
    cases["edge"] = edge
    ann_golden = {}
    base = RawGPT4oBaseline.__new__(RawGPT4oBaseline)
    for name, anns in cases.items():
        objs_ = sg._parse_annotations(anns)
        cats = sg._categorize_objects(objs_)
        zones = sg._build_spatial_zones(objs_)
        ann_golden[name] = {
            "annotations": anns,
            "parsed": [{k: (float(v) if k == "distance" else v) for k, v in o.items()} for o in objs_],
            "categorized": {k: [o["id"] for o in v] for k, v in cats.items()},
            "zones": {k: [o["id"] for o in v] for k, v in zones.items()},
            "describe": base._describe_annotations([a for a in anns if "category_name" in a]) if name != "edge" else None,
        }
    golden = {"annotations": ann_golden, "process_mock": process_golden,
              "get_direction": {"angles_deg": [float(a) for a in angles], "radius": 12.5, "labels": dirs},
              "describe_point_cloud_mock": base._describe_point_cloud(sample["point_cloud"]),
              "describe_point_cloud_empty": base._describe_point_cloud(np.zeros((0, 4), np.float32)),
              "docs_log_kat": {"scene": "docs_scene_1", "row": 2, "distance_rounded": 1186.3, "direction": "front",
                               "source": "docs/assets/scene_1_ca9a282c.log:821-832"}}
    def default(o):
        if isinstance(o, (np.floating,)): return float(o)
        if isinstance(o, (np.integer,)): return int(o)
        if isinstance(o, np.ndarray): return o.tolist()
        raise TypeError(type(o))
    json.dump(golden, open(os.path.join(HERE, "reference_golden.json"), "w"), indent=1, default=default)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
