"""The oracle (oracle/numpy_ref.py and oracle/c/msc_oracle.c) against golden vectors produced by the reference
itself (tests/golden/make_golden.py).  This is what pins the oracle for rows a3-a13 of SURVEY.md section 8."""
import math
import os

import numpy as np
import pytest

from oracle import numpy_ref as R
from tests import oracle_bridge as OB
from msc_geom.layout import GeomParams

CASES = ["mock", "edge", "synth", "empty", "flatobj"]
P800 = GeomParams(bev_res=800)


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, f"keyframe_{name}.npz"))


@pytest.mark.parametrize("name", CASES)
def test_numpy_filter_split_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    kept = R.preprocess_point_cloud(g["points"])
    ground, obj = R.segment_ground(kept)
    assert np.array_equal(kept, g["kept"]) and np.array_equal(ground, g["ground"]) and np.array_equal(obj, g["object"])


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_filter_split_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    pts = g["points"]
    if pts.shape[0] == 0:
        pytest.skip("empty cloud has no rows to index")
    k, gi, oi = OB.oracle_keyframe_filter_split(pts, P800)
    assert np.array_equal(pts[k], g["kept"]) and np.array_equal(pts[gi], g["ground"]) and np.array_equal(pts[oi], g["object"])


def test_c_oracle_filter_on_strided_view(golden_dir):
    """The real loader hands over a 20-byte-pitch view (nuscenes_loader.py:152-155)."""
    g = load(golden_dir, "synth")
    raw = np.zeros((g["points"].shape[0], 5), np.float32)
    raw[:, :4] = g["points"]
    k, gi, oi = OB.oracle_keyframe_filter_split(raw[:, :4], P800)
    assert np.array_equal(raw[k][:, :4], g["kept"]) and np.array_equal(raw[oi][:, :4], g["object"])


@pytest.mark.parametrize("name", CASES)
def test_numpy_bev_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    bev = R.generate_multi_layer_bev(g["ground"], g["object"])
    for k in ("semantic", "height", "density"):
        assert bev[k].dtype == g[k].dtype and np.array_equal(bev[k], g[k]), k


@pytest.mark.parametrize("name", ["mock", "edge", "synth", "flatobj"])
def test_c_oracle_bev_raster_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    pts = g["points"]
    _, gi, oi = OB.oracle_keyframe_filter_split(pts, P800)
    count, height, sem = OB.oracle_keyframe_bev(pts, gi, oi)
    bev = R.finish_bev(count.astype(np.int64), height, sem)
    for k in ("semantic", "height", "density"):
        assert np.array_equal(bev[k], g[k]), k


def test_cluster_metadata_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "clusters_synth.npz"))
    obj, labels = g["object"], g["labels"]
    order = [l for l in set(labels.tolist()) if l != -1 and (labels == l).sum() >= 5]  # lidar_agent.py:154-165
    assert len(order) == len(g["num_points"])
    aabb = OB.oracle_cluster_aabb(obj, labels, int(labels.max()) + 1)
    for i, l in enumerate(order):
        m = R.cluster_metadata(obj[labels == l])
        assert np.array_equal(m["center"], g["center"][i]) and np.array_equal(m["dimensions"], g["dimensions"][i])
        assert m["distance"] == g["distance"][i] and m["num_points"] == g["num_points"][i] and m["direction"] == str(g["direction"][i])
        assert np.array_equal(aabb[l, 6:9], g["center"][i]) and aabb[l, 9] == g["distance"][i] and aabb[l, 10] == g["num_points"][i]
        assert np.array_equal(aabb[l, 3:6] - aabb[l, 0:3], g["dimensions"][i])


def test_get_direction_matches_reference(golden_json):
    gd = golden_json["get_direction"]
    for a, lab in zip(gd["angles_deg"], gd["labels"]):
        p = np.array([np.cos(np.deg2rad(a)), np.sin(np.deg2rad(a))], np.float32) * np.float32(gd["radius"])
        assert R.get_direction(p) == lab


@pytest.mark.parametrize("case", ["mock", "docs_scene_1", "docs_scene_2", "docs_scene_3", "edge"])
def test_annotation_path_matches_reference(golden_json, case):
    g = golden_json["annotations"][case]
    anns = g["annotations"]
    parsed = R.parse_annotations(anns)
    assert len(parsed) == len(g["parsed"])
    for a, b in zip(parsed, g["parsed"]):
        assert a["id"] == b["id"] and a["category"] == b["category"] and a["direction"] == b["direction"]
        assert a["state"] == b["state"] and a["visibility"] == b["visibility"]
        assert (float(a["distance"]) == b["distance"]) or (math.isnan(a["distance"]) and math.isnan(b["distance"]))
    assert {k: [o["id"] for o in v] for k, v in R.categorize_objects(parsed).items()} == g["categorized"]
    assert {k: [o["id"] for o in v] for k, v in R.build_spatial_zones(parsed).items()} == g["zones"]
    # the C table (numeric half of the same functions)
    xy = np.array([a.get("translation", [0, 0, 0])[:2] for a in anns], np.float64)
    vel = np.zeros((len(anns), 2))
    for i, a in enumerate(anns):
        v = a.get("velocity", None)
        ok = isinstance(v, (list, tuple)) and len(v) >= 2 and v[0] is not None and v[1] is not None
        vel[i] = v[:2] if ok else [0.0, 0.0]
    t = OB.oracle_annotation_table(xy, vel)
    zone_names = [z[0] for z in R.SPATIAL_ZONES]
    zone_of = {oid: zn for zn, ids in g["zones"].items() for oid in ids}
    for i, b in enumerate(g["parsed"]):
        assert t["distance"][i] == b["distance"]
        assert R.DIRECTIONS_4[t["direction"][i]] == b["direction"]
        assert ("moving" if t["moving"][i] else "stopped") == b["state"]
        assert (zone_names[t["zone"][i]] if t["zone"][i] != 255 else None) == zone_of.get(b["id"])
    if g["describe"] is not None:
        assert R.describe_annotations(anns) == g["describe"]
        n = len(anns)
        assert f"- Front region: {int((t['region_bits'] & 1).sum())} objects" in g["describe"]
        assert f"- Right region: {n - int(((t['region_bits'] >> 1) & 1).sum())} objects" in g["describe"]


def test_docs_log_known_answer(golden_json):
    """docs/assets/scene_1_ca9a282c.log:821-832: vehicle.car at (353.794, 1132.355) -> 1186.3 m, 'front' (frame bug preserved)."""
    kat = golden_json["docs_log_kat"]
    a = golden_json["annotations"][kat["scene"]]["annotations"][kat["row"]]
    assert a["translation"][:2] == [353.794, 1132.355]
    p = R.parse_annotations([a])[0]
    assert round(float(p["distance"]), 1) == kat["distance_rounded"] and p["direction"] == kat["direction"]


def test_describe_point_cloud(golden_dir, golden_json):
    pts = load(golden_dir, "mock")["points"]
    assert R.describe_point_cloud(pts) == golden_json["describe_point_cloud_mock"]
    assert R.describe_point_cloud(np.zeros((0, 4), np.float32)) == golden_json["describe_point_cloud_empty"]
    mm, acc = OB.oracle_cloud_stats(pts)
    assert np.array_equal(mm[:3], pts[:, :3].min(0)) and np.array_equal(mm[3:], pts[:, :3].max(0))
    mean = np.sqrt(pts[:, 0] ** 2 + pts[:, 1] ** 2).mean()
    assert abs(acc / len(pts) - mean) <= 1e-5 * mean
    assert f"{acc / len(pts):.1f} m" in golden_json["describe_point_cloud_mock"]


def test_cluster_views_match_reference(golden_dir):
    """the NumPy restatement of _generate_cluster_visualization (lidar_agent.py:241-356) and of
    the batch mosaic (:366-386) against images produced by the reference itself."""
    g = np.load(os.path.join(golden_dir, "cluster_views.npz"))
    imgs = []
    for i in range(int(g["n"])):
        img = R.generate_cluster_visualization(g[f"pts_{i}"])
        assert np.array_equal(img, g[f"img_{i}"]), i
        imgs.append(img)
    assert np.array_equal(R.cluster_mosaic(imgs[:5]), g["mosaic"])
