"""Golden outputs of the reference's agent boundaries the mirrors must reproduce (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_agent_golden.py

  * SceneGraphAgent.process with the LLM call failing -> the reference's local fallback graph + its text summary
    (scenegraph_agent.py:148-178, 379-421, 423-490), and with a canned HierarchicalSceneGraph reply -> the summary of a full graph;
  * CameraAgent.process with canned replies -> result keys, the per-camera split (camera_agent.py:75-127) and the context text block
    it puts in front of the images (:42-47).
The Azure client is never touched: BaseAgent.call_llm is replaced.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/src")

from agents.base_agent import BaseAgent  # noqa: E402
from agents.content_transform import scenegraph_agent as sgm  # noqa: E402
from agents.content_transform.camera_agent import CameraAgent  # noqa: E402
from nuscenes_loader import create_loader  # noqa: E402


def main():
    np.random.seed(0)
    sample = create_loader(None, use_mock=True).get_sample_by_scene_index(0, 0)
    out = {}

    # ---- scene graph: LLM failure -> fallback graph
    def boom(self, messages, temperature=0.7, max_retries=8, response_format=None):
        raise RuntimeError("stubbed LLM failure")
    BaseAgent.call_llm = boom
    sg = sgm.SceneGraphAgent(object(), "m", "SceneGraphAgent")
    r = sg.process(sample["annotations"])
    out["scenegraph_fallback"] = {"annotations": sample["annotations"], "result": r}

    # ---- scene graph: a canned structured reply -> summary of a populated graph
    obj = dict(id="obj_0", category="car", subcategory="sedan", position_x=10.0, position_y=2.0, position_z=0.5, distance_to_ego=10.2,
               direction="right", state="moving", visibility="high")
    ped = dict(obj, id="obj_1", category="pedestrian", subcategory="adult")
    graph = sgm.HierarchicalSceneGraph(
        scene_summary="Two road users near the ego vehicle",
        environment=sgm.EnvironmentContext(lighting="day", weather="clear", visibility_overall="good", location_type="urban"),
        road_structure=sgm.RoadStructure(road_type="urban street", lanes=sgm.LaneInfo(lane_count=2, lane_type="urban", ego_lane_position="right",
                                                                                      lane_markings=["dashed"]),
                                         road_elements=[sgm.RoadElement(element_type="sign", description="stop sign", location="front right")],
                                         surface_condition="dry"),
        traffic_participants=sgm.TrafficParticipants(vehicles=[sgm.SceneObject(**obj)], cyclists=[], vulnerable_road_users=[sgm.SceneObject(**ped)]),
        sidewalk_areas=sgm.SidewalkArea(has_sidewalk=True, pedestrians=[sgm.SceneObject(**ped)], static_objects=[], location="right"),
        static_infrastructure=sgm.StaticInfrastructure(barriers=[sgm.SceneObject(**obj)], traffic_cones=[], construction=[sgm.SceneObject(**obj)], other=[]),
        spatial_zones=[sgm.SpatialZone(zone_name="right_medium", objects=[sgm.SceneObject(**obj)], is_clear=False, criticality="medium"),
                       sgm.SpatialZone(zone_name="front_close", objects=[], is_clear=True, criticality="low")],
        safety_critical_elements=["pedestrian near the kerb"], total_objects=2)
    BaseAgent.call_llm = lambda self, messages, temperature=0.7, max_retries=8, response_format=None: graph
    r2 = sg.process(sample["annotations"])
    out["scenegraph_llm"] = {"result": r2}

    # ---- camera agent: canned replies
    cams = sample["camera_names"]
    imgs = [np.full((6, 8, 3), 10 * i, np.uint8) for i in range(len(cams))]
    captured = {}
    replies = {
        "sectioned": "\n".join(f"{c}: view {i} shows a road." for i, c in enumerate(cams)),
        "spaces": "cam front: a car ahead. Cam Back Left: nothing. The cam_back view is empty.",
        "unsectioned": "A generic description without any camera label.",
    }
    cam = CameraAgent(object(), "m", "CameraAgent")
    ctx = {"box_projection": {"CAM_FRONT": {"visible_objects": 1, "objects": [{"annotation": 0, "category": "vehicle.car", "bbox": [1.0, 2.0, 3.5, 4.0]}]}}}
    res = {}
    for name, reply in replies.items():
        def fake(self, messages, temperature=0.7, max_retries=8, response_format=None, _r=reply):
            captured["messages"] = messages
            return _r
        BaseAgent.call_llm = fake
        res[name] = cam.process(imgs, cams, context=ctx if name == "sectioned" else None)
        if name == "sectioned":
            content = captured["messages"][1]["content"]
            out["camera_context_text"] = content[0]["text"]
            out["camera_prompt_text"] = content[1]["text"]
            out["camera_first_image_url"] = content[3]["image_url"]["url"]
    out["camera"] = {"camera_names": cams, "replies": replies, "context": ctx, "results": res}
    json.dump(out, open(os.path.join(HERE, "agent_golden.json"), "w"), indent=1, default=lambda o: o.tolist() if isinstance(o, np.ndarray) else float(o))
    print("wrote agent_golden.json", {k: type(v).__name__ for k, v in out.items()})


if __name__ == "__main__":
    main()
