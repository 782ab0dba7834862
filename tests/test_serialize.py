"""the strings built from the evidence equal the reference's, character for character
(goldens made by running the reference: tests/golden/make_prompt_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from msc_geom import serialize as S


@pytest.fixture(scope="module")
def prompt_golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "prompt_golden.json")))


def _from_ids(case):
    objs = {o["id"]: o for o in case["parsed"]}
    cats = {k: [objs[i] for i in v] for k, v in case["categorized"].items()}
    zones = {k: [objs[i] for i in v] for k, v in case["zones"].items()}
    return cats, zones


@pytest.mark.parametrize("case", ["mock", "docs_scene_1", "docs_scene_2", "docs_scene_3", "edge"])
def test_scene_graph_prompt_equals_reference(golden_json, prompt_golden, case):
    c = golden_json["annotations"][case]
    cats, zones = _from_ids(c)
    want = prompt_golden["scene_graph_prompts"][case]
    assert S.scene_graph_user_prompt(cats, zones, c["annotations"], want["context"]) == want["user_prompt"]


def test_image_to_base64_equals_reference(prompt_golden):
    for name, v in prompt_golden["image_to_base64"].items():
        img = np.array(v["data"], dtype=v["dtype"]).reshape(v["shape"])
        assert S.image_to_base64(img) == v["base64"], name
    big = np.random.default_rng(7)
    for shape, dt in (((12, 9, 3), np.uint8), ((5, 7), np.uint8)):  # replay the generator's draws to reach the 800x800 image
        big.integers(0, 256, shape, dtype=dt)
    big.random((6, 4, 3))
    img = big.integers(0, 256, (800, 800, 3), dtype=np.uint8)
    assert hashlib.sha256(S.image_to_base64(img).encode()).hexdigest() == prompt_golden["image_to_base64_800"]["sha256"]


def test_ext_evidence_block_is_json_and_consistent():
    anns = [{"category_name": "vehicle.car"}, {"category_name": "human.pedestrian.adult"}, {"category_name": "movable_object.barrier"}]
    cnt = np.array([12, 0, 3], np.uint32)
    near = np.array([7.26, np.inf, 21.04], np.float32)
    cen = np.array([[7.3, 0.2, -0.5], [0, 0, 0], [20.9, 3.3, 0.1]], np.float32)
    vis = np.array([[1, 0], [0, 0], [1, 1]], np.uint8)
    rel = {"dist": np.array([[0, 5, 14], [5, 0, 9.5], [14, 9.5, 0]], np.float32), "category": np.array([[0, 1, 2], [3, 0, 0], [2, 1, 0]], np.uint8),
           "overlap": np.zeros((3, 3), np.uint8), "labels": ("ahead", "left", "behind", "right")}
    ev = S.ext_evidence(anns, cnt, near, cen, vis, ["CAM_FRONT", "CAM_BACK"], relations=rel, max_pairs=4)
    txt = json.dumps(ev)
    back = json.loads(txt)
    assert [o["lidar_points"] for o in back["objects"]] == [12, 0, 3]
    assert back["objects"][1]["nearest_point_m"] is None and back["objects"][1]["visible_in"] == []
    assert back["objects"][2]["visible_in"] == ["CAM_FRONT", "CAM_BACK"] and back["objects"][0]["nearest_point_m"] == 7.3
    assert [p["distance_m"] for p in back["nearest_pairs"]] == [5.0, 5.0, 9.5, 9.5] and back["nearest_pairs"][0]["relation"] == "left"


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["mock", "docs_scene_1", "edge"])
def test_prompt_through_the_gpu_annotation_path(engine, golden_json, prompt_golden, case):
    """annotations -> CUDA annotation table -> categories / zones -> the reference's prompt string."""
    from msc_geom.scenegraph_agent import SceneGraphAgent
    agent = SceneGraphAgent(None, "m", "SceneGraphAgent", engine=engine)
    c = golden_json["annotations"][case]
    want = prompt_golden["scene_graph_prompts"][case]
    assert agent.scene_graph_prompt(c["annotations"], want["context"]) == want["user_prompt"]


@pytest.mark.gpu
def test_ext_evidence_from_a_fused_run(engine):
    from msc_geom.synthetic import CAMERA_CHANNELS
    from msc_geom.layout import pack_batch
    from msc_geom.synthetic import make_sample
    from tests import oracle_bridge as OB
    from msc_geom.layout import GeomParams
    s = make_sample(77, n_sweeps=2, n_boxes=12)
    hb = pack_batch([s])
    import torch
    out = engine.run_fused(engine.upload(hb)); torch.cuda.synchronize()
    got = out.to_host(with_bev=False)
    ev = S.ext_evidence(s["annotations"], got["box_count"], got["box_nearest"], got["box_centroid"], got["proj_visible"], CAMERA_CHANNELS)
    ref = OB.oracle_fused(hb, 0, GeomParams())
    assert [o["lidar_points"] for o in ev["objects"]] == [int(v) for v in ref["box_count"]]
    assert json.loads(json.dumps(ev)) == ev
