"""Golden vectors for SURVEY.md section 8(f) rank 4 (evidence serialisation into prompts), made by IMPORTING THE REFERENCE
(/root/reference/src, read-only; build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_prompt_golden.py

call_llm is replaced by a recorder, so the user prompt SceneGraphAgent._generate_scene_graph builds
(scenegraph_agent.py:327-366) is captured exactly as the reference would send it; the Azure client is never touched.
LiDARAgent._image_to_base64 (lidar_agent.py:819-832) is run on small seeded images.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/src")

from agents.base_agent import BaseAgent  # noqa: E402
from agents.content_transform.lidar_agent import LiDARAgent  # noqa: E402
from agents.content_transform.scenegraph_agent import SceneGraphAgent  # noqa: E402


def main():
    ref = json.load(open(os.path.join(HERE, "reference_golden.json")))
    captured = []

    def recorder(self, messages, temperature=0.7, max_retries=8, response_format=None):
        captured.append(messages)
        raise RuntimeError("recorded")  # the reference falls back to its minimal graph (scenegraph_agent.py:379-421)

    BaseAgent.call_llm = recorder
    sg = SceneGraphAgent(object(), "m", "SceneGraphAgent")
    contexts = {"mock": None, "docs_scene_1": {"lidar": {"num_objects": 21, "nearest_m": 3.25}, "camera": ["CAM_FRONT: two cars"]},
                "docs_scene_2": {"note": "x" * 900}, "docs_scene_3": None, "edge": {}}
    prompts = {}
    for name, case in ref["annotations"].items():
        captured.clear()
        anns = case["annotations"]
        objs = sg._parse_annotations(anns)
        sg._generate_scene_graph(sg._categorize_objects(objs), sg._build_spatial_zones(objs), anns, contexts.get(name))
        assert len(captured) == 1
        prompts[name] = {"context": contexts.get(name), "user_prompt": captured[0][1]["content"]}
    rng = np.random.default_rng(7)
    images = {"bgr_u8": rng.integers(0, 256, (12, 9, 3), dtype=np.uint8), "gray_u8": rng.integers(0, 256, (5, 7), dtype=np.uint8),
              "float01": rng.random((6, 4, 3)).astype(np.float32), "black": np.zeros((3, 3, 3), np.uint8)}
    b64 = {k: {"shape": list(v.shape), "dtype": str(v.dtype), "data": v.tolist(), "base64": LiDARAgent._image_to_base64(v)} for k, v in images.items()}
    big = rng.integers(0, 256, (800, 800, 3), dtype=np.uint8)
    b64_big = {"seed": 7, "shape": [800, 800, 3], "sha256": hashlib.sha256(LiDARAgent._image_to_base64(big).encode()).hexdigest()}
    json.dump({"scene_graph_prompts": prompts, "image_to_base64": b64, "image_to_base64_800": b64_big},
              open(os.path.join(HERE, "prompt_golden.json"), "w"), indent=1)
    print("wrote prompt_golden.json", {k: len(v["user_prompt"]) for k, v in prompts.items()})


if __name__ == "__main__":
    main()
