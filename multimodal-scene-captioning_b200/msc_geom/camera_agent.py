"""Camera-side evidence.  The reference's CameraAgent does no geometry (camera_agent.py:12-127 is base64 + one VLM call);
its `process(images, camera_names, context=None)` JSON-dumps `context` into the prompt (:43-47), which is the natural carrier
for the [EXT] box -> camera projection evidence built here (devkit get_sample_data + view_points + box_in_image, App. A.3)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from . import ops
from .engine import GeometryEngine
from .layout import boxes_from_annotations


def projection_evidence(engine: GeometryEngine, sample: Dict[str, Any], image_size=(1600, 900)) -> Dict[str, Any]:
    """Per camera: which annotations are visible and their clipped 2D extents.  Needs the additive `cameras` key
    (ego_pose, calib, intrinsic per camera); returns {} for plain reference-style samples."""
    cams = sample.get("cameras") or []
    anns = sample.get("annotations") or []
    if not cams or not anns:
        return {}
    boxes = boxes_from_annotations(anns)
    pose = np.stack([np.asarray(c.get("ego_pose", sample.get("ego_pose")), np.float64) for c in cams])
    calib = np.stack([np.asarray(c["calib"], np.float64) for c in cams])
    K = np.stack([np.asarray(c["intrinsic"], np.float64).reshape(9) for c in cams])
    vis, ext = ops.project_boxes(engine, boxes, pose, calib, K, image_size[0], image_size[1])
    out: Dict[str, Any] = {}
    for ci, c in enumerate(cams):
        rows = [{"annotation": i, "category": anns[i].get("category_name", "unknown"), "bbox": [round(float(v), 1) for v in ext[i, ci]]}
                for i in np.nonzero(vis[:, ci])[0]]
        out[c.get("channel", f"CAM_{ci}")] = {"visible_objects": len(rows), "objects": rows}
    return out
