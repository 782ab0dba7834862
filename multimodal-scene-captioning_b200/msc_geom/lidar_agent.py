"""LiDARAgent -- host-side mirror of the reference agent's LOCAL half
(/root/reference/src/agents/content_transform/lidar_agent.py), with the numeric work on the GPU.

Same constructor, attribute names, method names, argument meaning and return shapes as the reference
(SURVEY.md section 8(b)), so the three geometry methods the reference's scripts call directly
(src/export_sample_data.py:59-65, src/generate_detailed_logs.py:213-215) and `process()` are drop-ins.
What stays on the host, and why:
  * the cluster LIST ORDER follows CPython set iteration over the labels, like the reference (lidar_agent.py:154-159);
    DBSCAN itself runs on the device with scikit-learn's exact labelling (csrc/dbscan.cu).
  * the cv2 overlays drawn after the raster (lidar_agent.py:599-634) and the log1p density normalisation, which is
    evaluated with NumPy's own float32 log1p through a count -> value table so the uint8 layer is bit-identical.
  * every remote LLM call: injected as callables (`llm`, `cluster_classifier`); absent callables yield the
    reference's own fallbacks ('unknown' clusters are dropped, lidar_agent.py:228).
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

from . import ops
from .engine import GeometryEngine
from .layout import GeomParams

DIRECTION_LABELS = ("front_right", "front", "front_left", "left", "back_left", "back", "back_right", "right")


@dataclass
class DetectedObject:
    """Same fields as the reference dataclass (lidar_agent.py:18-27)."""
    category: str
    position: np.ndarray
    dimensions: np.ndarray
    num_points: int
    distance: float
    direction: str
    confidence: float


def direction_label(position_2d) -> str:
    """8-way label of lidar_agent.py:506-530: 45-degree bins starting at 337.5, in the reference's (rotated) label order."""
    x, y = position_2d
    angle = (np.arctan2(y, x) * 180 / np.pi + 360) % 360
    if angle >= 337.5 or angle < 22.5:
        return DIRECTION_LABELS[0]
    lo = 22.5
    for label in DIRECTION_LABELS[1:7]:
        if lo <= angle < lo + 45.0:
            return label
        lo += 45.0
    return DIRECTION_LABELS[7]


def finish_bev_layers(count: np.ndarray, height: np.ndarray, semantic: np.ndarray, res: int, bev_range) -> Dict[str, np.ndarray]:
    """Host half of _generate_multi_layer_bev: density normalisation (lidar_agent.py:563-564), ego cross, vertical
    flips, range rings and labels (:599-634).  `count`/`height`/`semantic` are the unflipped GPU raster layers."""
    import cv2
    cmax = int(count.max()) if count.size else 0
    lut = np.log1p(np.arange(cmax + 1, dtype=np.float32))  # NumPy's float32 log1p, one value per distinct count
    dens = lut[count.astype(np.int64)]
    density = (dens / dens.max() * 255).astype(np.uint8) if cmax > 0 else dens.astype(np.uint8)
    vis = np.ascontiguousarray(semantic)
    c, arm = res // 2, 15
    green = (0, 255, 0)
    cv2.line(vis, (c - arm, c), (c + arm, c), green, 3)
    cv2.line(vis, (c, c - arm), (c, c + arm), green, 3)
    vis, height, density = (cv2.flip(np.ascontiguousarray(a), 0) for a in (vis, height, density))
    for metres in (10, 20, 30, 40):
        radius = int(metres / (2 * bev_range) * res)
        cv2.circle(vis, (c, c), radius, (100, 100, 100), 1)
        cv2.putText(vis, f"{metres}m", (c + 5, c - radius + 15), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (150, 150, 150), 1)
    for text, org in (("FRONT", (c - 25, 20)), ("BACK", (c - 20, res - 10)), ("L", (10, c + 5)), ("R", (res - 20, c + 5))):
        cv2.putText(vis, text, org, cv2.FONT_HERSHEY_SIMPLEX, 0.6, (200, 200, 200), 2)
    return {"semantic": vis, "height": height, "density": density}


def finish_cluster_view(grid: np.ndarray, img_size: int = 256) -> np.ndarray:
    """Host half of _generate_cluster_visualization: axes and titles (lidar_agent.py:299-313, :353-354) on top of the GPU
    disc raster.  They sit well inside their quadrant, a disc bleeds at most two pixels over a quadrant border, so drawing
    them after all discs gives the reference's image."""
    import cv2
    grid = np.ascontiguousarray(grid)
    half = img_size // 2
    for qx, qy, title in ((0, 0, "Top (XY)"), (1, 0, "Side (XZ)"), (0, 1, "Front (YZ)"), (1, 1, "3D View")):
        ox, oy = qx * img_size, qy * img_size
        if (qx, qy) != (1, 1):
            cv2.line(grid, (ox + half, oy + half), (ox + half + 30, oy + half), (0, 0, 255), 2)   # x axis, red
            cv2.line(grid, (ox + half, oy + half), (ox + half, oy + half - 30), (0, 255, 0), 2)   # y axis, green
        cv2.putText(grid, title, (ox + 10, oy + 20), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (0, 0, 0), 1)
    return grid


def cluster_mosaic(images: List[np.ndarray]) -> np.ndarray:
    """The batch image of _classify_batch_with_llm (lidar_agent.py:366-386): up to three columns, '#idx' labels."""
    import cv2
    if len(images) == 1:
        return images[0]
    cols = min(3, len(images))
    rows = (len(images) + cols - 1) // cols
    h, w = images[0].shape[:2]
    sheet = np.full((rows * h, cols * w, 3), 255, dtype=np.uint8)
    for idx, img in enumerate(images):
        r, c = divmod(idx, cols)
        sheet[r * h:(r + 1) * h, c * w:(c + 1) * w] = img
        cv2.putText(sheet, f"#{idx}", (c * w + 10, r * h + 50), cv2.FONT_HERSHEY_SIMPLEX, 1.5, (255, 0, 0), 3)
    return sheet


class LiDARAgent:
    def __init__(self, client, model: str, agent_name: str, engine: Optional[GeometryEngine] = None,
                 llm: Optional[Callable[..., str]] = None, cluster_classifier: Optional[Callable[..., List[dict]]] = None):
        self.client, self.model, self.agent_name = client, model, agent_name
        self.dbscan_eps = 0.5          # lidar_agent.py:44
        self.dbscan_min_samples = 10   # :45
        self.bev_resolution = 800      # :48
        self.bev_range = 50            # :49
        self.engine = engine or GeometryEngine()
        self.llm = llm
        self.cluster_classifier = cluster_classifier

    # ------------------------------------------------------------------ remote half (injected)
    def call_llm(self, messages, temperature: float = 0.7, **kw) -> str:
        if self.llm is None:
            return ""
        return self.llm(messages, temperature=temperature, **kw)

    # ------------------------------------------------------------------ local geometry (GPU)
    def _params(self, ground_threshold: float = -1.4) -> GeomParams:
        return GeomParams(range_max=float(self.bev_range), bev_range=float(self.bev_range), bev_res=int(self.bev_resolution),
                          ground_z=float(ground_threshold))

    def _preprocess_point_cloud(self, pc: np.ndarray) -> np.ndarray:
        return ops.keyframe_filter_split(self.engine, pc, self._params())[0]

    def _segment_ground(self, pc: np.ndarray, ground_threshold: float = -1.4) -> Tuple[np.ndarray, np.ndarray]:
        _, ground, obj = ops.keyframe_filter_split(self.engine, pc, self._params(ground_threshold), split_only=True)
        return ground, obj

    def _generate_multi_layer_bev(self, ground_points: np.ndarray, object_points: np.ndarray) -> Dict[str, np.ndarray]:
        res, r = int(self.bev_resolution), self.bev_range
        count, height, sem = ops.keyframe_bev_layers(self.engine, ground_points, object_points, res, float(r))
        return finish_bev_layers(count, height, sem, res, r)

    def _generate_cluster_visualization(self, points: np.ndarray, img_size: int = 256) -> np.ndarray:
        """One cluster -> 512x512 4-view image (lidar_agent.py:241-356)."""
        if img_size != 256:
            raise ValueError("the device raster is built for the reference's img_size = 256")
        return finish_cluster_view(ops.cluster_views(self.engine, points, [np.arange(len(points))])[0])

    def _cluster_visualizations(self, object_points: np.ndarray, labels: np.ndarray, order: List[int]) -> List[np.ndarray]:
        """All clusters of a cloud in one launch (the reference loops over clusters, lidar_agent.py:198-209)."""
        discs = ops.cluster_views(self.engine, object_points, [np.nonzero(labels == l)[0] for l in order])
        return [finish_cluster_view(d) for d in discs]

    def _cluster_metadata(self, object_points: np.ndarray, labels: np.ndarray, order: List[int]) -> List[dict]:
        n_clusters = int(labels.max()) + 1 if labels.size else 0
        table = ops.cluster_aabb(self.engine, object_points, labels, n_clusters)
        meta = []
        for i, lab in enumerate(order):
            row = table[lab]
            mn, mx, center = row[0:3], row[3:6], row[6:9]
            meta.append({"index": i, "center": center.copy(), "dimensions": mx - mn, "distance": row[9], "direction": direction_label(center[:2]),
                         "num_points": int(row[10])})
        return meta

    def _detect_objects_3d(self, object_points: np.ndarray) -> List[DetectedObject]:
        """lidar_agent.py:134-239 with the local work on the device: DBSCAN (scikit-learn's labelling), per-cluster boxes, and the
        4-view images of ALL clusters in one launch; the classifier sees what the reference's sees -- per batch of ten, the cluster
        images and their metadata (`_classify_batch_with_llm(cluster_images, cluster_metadata)`, :358)."""
        if len(object_points) < self.dbscan_min_samples:
            return []
        labels = ops.dbscan(self.engine, object_points, self.dbscan_eps, self.dbscan_min_samples)
        uniq = set(labels)
        uniq.discard(-1)
        order = [int(l) for l in uniq if int((labels == l).sum()) >= 5]  # set-iteration order, like lidar_agent.py:154-165
        if not order:
            return []
        labels = labels.astype(np.int32)
        meta = self._cluster_metadata(object_points, labels, order)
        images = self._cluster_visualizations(object_points, labels, order)
        out: List[DetectedObject] = []
        for start in range(0, len(meta), 10):  # the reference classifies in batches of 10 (:189)
            batch = meta[start:start + 10]
            cls = self._classify_batch_with_llm(images[start:start + 10], batch)
            for m, c in zip(batch, cls):
                if c["category"] != "unknown" and c["confidence"] > 0.3:
                    out.append(DetectedObject(c["category"], m["center"], m["dimensions"], m["num_points"], m["distance"], m["direction"],
                                              c["confidence"]))
        return out

    def _classify_batch_with_llm(self, cluster_images: List[np.ndarray], cluster_metadata: List[dict]) -> List[dict]:
        """The remote half of lidar_agent.py:358-504, injected: `cluster_classifier(images, mosaic, metadata)` gets the batch's cluster
        images, the mosaic the reference would send (:366-386) and the metadata.  Without one every cluster takes the reference's
        parse-failure default (:503)."""
        if self.cluster_classifier is None:
            return [{"category": "unknown", "confidence": 0.5} for _ in cluster_metadata]
        return self.cluster_classifier(cluster_images, cluster_mosaic(cluster_images), cluster_metadata)

    # ------------------------------------------------------------------ evidence -> features / report (host, tiny)
    def _extract_semantic_features(self, detected_objects: List[DetectedObject], ground_points: np.ndarray, object_points: np.ndarray) -> Dict[str, Any]:
        counts: Dict[str, int] = {}
        by_dir = {d: 0 for d in ("front", "back", "left", "right", "front_left", "front_right", "back_left", "back_right")}
        close = medium = far = vehicles = 0
        for o in detected_objects:
            counts[o.category] = counts.get(o.category, 0) + 1
            by_dir[o.direction] += 1
            if o.distance < 10:
                close += 1
            elif o.distance < 30:
                medium += 1
            else:
                far += 1
            vehicles += o.category in ("car", "truck", "bus")
        total = len(ground_points) + len(object_points)
        ratio = len(object_points) / total if total > 0 else 0
        return {
            "total_objects": len(detected_objects), "object_counts": counts, "objects_by_direction": by_dir,
            "distance_distribution": {"close": close, "medium": medium, "far": far},
            "scene_characteristics": {"object_point_ratio": float(ratio),
                                      "traffic_density": "heavy" if vehicles > 10 else "moderate" if vehicles > 5 else "light",
                                      "total_points": total},
            "nearest_object": min(detected_objects, key=lambda o: o.distance) if detected_objects else None,
        }

    def _generate_structured_report(self, semantic_features: Dict[str, Any], detected_objects: List[DetectedObject]) -> str:
        f = semantic_features
        lines = ["=== LiDAR Scene Analysis ===\n", f"Total detected objects: {f['total_objects']}"]
        if f["object_counts"]:
            lines.append("\nObject Distribution:")
            lines += [f"  - {n} {cat}(s)" for cat, n in sorted(f["object_counts"].items())]
        lines.append("\nSpatial Distribution:")
        for direction, n in f["objects_by_direction"].items():
            if n > 0:
                cats = ", ".join(set(o.category for o in detected_objects if o.direction == direction))
                lines.append(f"  - {direction}: {n} objects ({cats})")
        d = f["distance_distribution"]
        lines += ["\nDistance Distribution:", f"  - Close (<10m): {d['close']} objects", f"  - Medium (10-30m): {d['medium']} objects",
                  f"  - Far (>30m): {d['far']} objects"]
        near = f["nearest_object"]
        if near:
            lines += ["\nNearest Object:", f"  - Type: {near.category}", f"  - Distance: {near.distance:.1f}m", f"  - Direction: {near.direction}"]
        sc = f["scene_characteristics"]
        lines += ["\nScene Characteristics:", f"  - Traffic density: {sc['traffic_density']}", f"  - Object point ratio: {sc['object_point_ratio']:.2%}"]
        return "\n".join(lines)

    @staticmethod
    def _object_to_dict(obj: DetectedObject) -> Dict[str, Any]:
        return {"category": obj.category, "position": obj.position.tolist(), "dimensions": obj.dimensions.tolist(), "num_points": obj.num_points,
                "distance": float(obj.distance), "direction": obj.direction, "confidence": float(obj.confidence)}

    # ------------------------------------------------------------------ entry point (lidar_agent.py:51-101)
    def process(self, point_cloud: np.ndarray, context: Optional[Dict] = None) -> Dict[str, Any]:
        kept, ground, obj = ops.keyframe_filter_split(self.engine, point_cloud, self._params())
        detected = self._detect_objects_3d(obj)
        bev = self._generate_multi_layer_bev(ground, obj)
        features = self._extract_semantic_features(detected, ground, obj)
        report = self._generate_structured_report(features, detected)
        observations = self._scene_interpretation(report, bev, context)
        return {"agent": self.agent_name, "modality": "lidar", "detected_objects": [self._object_to_dict(o) for o in detected],
                "semantic_features": features, "structured_report": report, "observations": observations,
                "bev_metadata": {"num_objects": len(detected), "ground_points": len(ground), "object_points": len(obj)}}

    def _scene_interpretation(self, report: str, bev: Dict[str, np.ndarray], context: Optional[Dict]) -> str:
        if self.llm is None:
            return ""
        prompt = "Analyze this driving scene from LiDAR data:\n\n" + report
        if context:
            prompt += "\n\nAdditional context from other sensors:\n" + json.dumps(context, indent=2)
        return self.call_llm([{"role": "user", "content": prompt}], temperature=0.4, bev_semantic=bev["semantic"])


def create_lidar_agent(client, model: str, **kw) -> LiDARAgent:
    return LiDARAgent(client, model, "LiDARAgent", **kw)
