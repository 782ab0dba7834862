"""SceneGraphAgent -- host-side mirror of the reference agent's LOCAL half
(/root/reference/src/agents/content_transform/scenegraph_agent.py:127-295) plus the [EXT] pairwise relation table.

The numeric columns (distance, 4-way direction, moving flag, zone, region bits) come from one small CUDA kernel per call;
the string columns (category prefix stripping, visibility bucket, ids) stay on the host.  The reference's frame bug is
preserved: `translation` is used exactly as the loader hands it over (global frame), so existing keys
keep their reference values; ego-frame relations are additive under the new `relations()` method.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from . import ops, serialize
from .engine import GeometryEngine
from .layout import boxes_from_annotations

DIRECTIONS = ("front", "left", "back", "right")
ZONE_NAMES = ("front_close", "front_medium", "front_far", "left_close", "left_medium", "right_close", "right_medium", "back_close", "back_medium")
_PREFIXES = ("vehicle.", "human.pedestrian.", "movable_object.", "static_object.")
_CATEGORY_ROUTES = (("vehicles", ("car", "truck", "bus", "trailer")), ("cyclists", ("bicycle", "motorcycle")),
                    ("pedestrians", ("pedestrian", "adult", "child")), ("barriers", ("barrier",)), ("traffic_cones", ("cone",)),
                    ("construction", ("construction",)))


def _velocity_pair(v) -> Optional[tuple]:
    """The reference's acceptance rules for `velocity` (scenegraph_agent.py:209-225); None means 'stopped'."""
    if isinstance(v, (list, tuple)) and len(v) >= 2 and v[0] is not None and v[1] is not None:
        try:
            return float(v[0]), float(v[1])
        except (TypeError, ValueError):
            return None
    return None


class SceneGraphAgent:
    def __init__(self, client, model: str, agent_name: str, engine: Optional[GeometryEngine] = None):
        self.client, self.model, self.agent_name = client, model, agent_name
        self.engine = engine or GeometryEngine()
        # scenegraph_agent.py:136-146 (kept as an attribute like the reference; the kernel holds the same table)
        self.spatial_zones = {"front_close": (0, 10, "front"), "front_medium": (10, 30, "front"), "front_far": (30, 50, "front"),
                              "left_close": (0, 10, "left"), "left_medium": (10, 30, "left"), "right_close": (0, 10, "right"),
                              "right_medium": (10, 30, "right"), "back_close": (0, 10, "back"), "back_medium": (10, 30, "back")}
        self._last_table: Optional[Dict[str, np.ndarray]] = None
        self.scene_graph_fn = None  # the remote half: user prompt -> scene-graph dict (None: the reference's local fallback graph)

    def _table(self, annotations: List[Dict]) -> Dict[str, np.ndarray]:
        n = len(annotations)
        xy = np.zeros((n, 2))
        vel = np.zeros((n, 2))
        for i, a in enumerate(annotations):
            pos = a.get("translation", [0, 0, 0])
            xy[i] = (pos[0], pos[1])
            v = _velocity_pair(a.get("velocity", None))
            if v is not None:
                vel[i] = v
        return ops.annotation_table(self.engine, xy, vel)

    def _parse_annotations(self, annotations: List[Dict]) -> List[Dict]:
        t = self._table(annotations)
        self._last_table = t
        objects = []
        for i, a in enumerate(annotations):
            category = a.get("category_name", "unknown").lower()
            for p in _PREFIXES:
                category = category.replace(p, "")
            vis = str(a.get("visibility_token", ""))
            visibility = "high" if ("80" in vis or "100" in vis) else "medium" if ("40" in vis or "60" in vis) else "low"
            objects.append({"id": f"obj_{i}", "category": category, "position": a.get("translation", [0, 0, 0]), "distance": t["distance"][i],
                            "direction": DIRECTIONS[int(t["direction"][i])], "state": "moving" if t["moving"][i] else "stopped",
                            "visibility": visibility, "attributes": a.get("attribute_tokens", [])})
        return objects

    def _categorize_objects(self, objects: List[Dict]) -> Dict[str, List[Dict]]:
        out: Dict[str, List[Dict]] = {name: [] for name, _ in _CATEGORY_ROUTES}
        out["other"] = []
        for obj in objects:
            cat = obj["category"]
            out[next((name for name, keys in _CATEGORY_ROUTES if any(k in cat for k in keys)), "other")].append(obj)
        return out

    def _build_spatial_zones(self, objects: List[Dict]) -> Dict[str, List[Dict]]:
        zones: Dict[str, List[Dict]] = {name: [] for name in self.spatial_zones}
        t = self._last_table
        if t is None or len(t["zone"]) != len(objects) or any(o["distance"] != t["distance"][i] for i, o in enumerate(objects)):
            # objects did not come from the last _parse_annotations call: classify their positions on the device again
            t = ops.annotation_table(self.engine, np.array([o["position"][:2] for o in objects], dtype=np.float64).reshape(-1, 2),
                                     np.zeros((len(objects), 2)))
        for i, obj in enumerate(objects):
            z = int(t["zone"][i])
            if z != 255:
                zones[ZONE_NAMES[z]].append(obj)
        return zones

    def region_counts(self, annotations: List[Dict]) -> Dict[str, int]:
        """RawGPT4oBaseline._describe_annotations region counts (baseline_gpt4o.py:304-317)."""
        t = self._table(annotations)
        n = len(annotations)
        front, left = int((t["region_bits"] & 1).sum()), int(((t["region_bits"] >> 1) & 1).sum())
        return {"front": front, "back": n - front, "left": left, "right": n - left}

    def relations(self, annotations: List[Dict], ego_pose=None) -> Dict[str, Any]:
        """[EXT] pairwise relation table: distance, bearing of j from i, ahead/left/behind/right with the bins of
        scenegraph_agent.py:194-201, and BEV footprint overlap.  With `ego_pose` the table is in the ego frame."""
        rel = ops.relation_table(self.engine, boxes_from_annotations(annotations), None if ego_pose is None else np.asarray(ego_pose, np.float64))
        rel["labels"] = ("ahead", "left", "behind", "right")
        return rel

    def scene_graph_prompt(self, annotations: List[Dict], context: Optional[Dict] = None) -> str:
        """The user message the reference sends to its LLM in _generate_scene_graph (scenegraph_agent.py:327-366), built from the
        GPU-computed annotation table; equal to the reference's string (tests/test_serialize.py)."""
        objs = self._parse_annotations(annotations)
        return serialize.scene_graph_user_prompt(self._categorize_objects(objs), self._build_spatial_zones(objs), annotations, context)

    # ------------------------------------------------------------------ entry point (scenegraph_agent.py:148-178)
    def process(self, annotations: List[Dict], context: Optional[Dict] = None) -> Dict[str, Any]:
        """Same keys as the reference: `agent`, `modality`, `scene_graph` (the dict form of its HierarchicalSceneGraph), `observations`
        (its text summary).  The graph itself is the reference's LLM half: `scene_graph_fn(user_prompt) -> dict` supplies it when
        injected; without one, or when it raises, the graph is the reference's own local fallback (:379-421).  The GPU-computed
        evidence rides along under the additive `evidence` key."""
        objs = self._parse_annotations(annotations)
        cats = self._categorize_objects(objs)
        zones = self._build_spatial_zones(objs)
        graph = self._generate_scene_graph(cats, zones, annotations, context)
        return {"agent": self.agent_name, "modality": "scene_graph", "scene_graph": graph, "observations": self._generate_summary(graph),
                "evidence": {"objects": objs, "categorized": {k: [o["id"] for o in v] for k, v in cats.items()},
                             "spatial_zones": {k: [o["id"] for o in v] for k, v in zones.items()}}}

    def _generate_scene_graph(self, categorized: Dict, spatial_zones: Dict, annotations: List[Dict], context: Optional[Dict]) -> Dict[str, Any]:
        fn = getattr(self, "scene_graph_fn", None)
        if fn is not None:
            try:
                return dict(fn(serialize.scene_graph_user_prompt(categorized, spatial_zones, annotations, context)))
            except Exception as e:  # noqa: BLE001 -- the reference catches everything here and falls back (:378-380)
                print(f"  \u26a0\ufe0f  Error generating scene graph: {e}")
        return fallback_scene_graph(len(annotations))

    def _generate_summary(self, scene_graph: Dict[str, Any]) -> str:
        """Text summary of a scene-graph dict, line for line the reference's (:423-490)."""
        graph = scene_graph
        env, road, lanes = graph["environment"], graph["road_structure"], graph["road_structure"]["lanes"]
        lines = ["=== Hierarchical Scene Graph ===\n", f"Scene: {graph['scene_summary']}", f"Total objects: {graph['total_objects']}\n",
                 "Environment:", f"  - Lighting: {env['lighting']}", f"  - Weather: {env['weather']}", f"  - Location: {env['location_type']}\n",
                 "Road Structure:", f"  - Type: {road['road_type']}", f"  - Lanes: {lanes['lane_count']} {lanes['lane_type']} lanes",
                 f"  - Ego position: {lanes['ego_lane_position']} lane"]
        if road["road_elements"]:
            lines.append(f"  - Elements: {len(road['road_elements'])} road signs/markings\n")
        tp = graph["traffic_participants"]
        lines += ["Traffic Participants:", f"  - Vehicles: {len(tp['vehicles'])}", f"  - Cyclists: {len(tp['cyclists'])}",
                  f"  - Vulnerable road users: {len(tp['vulnerable_road_users'])}\n"]
        sw = graph["sidewalk_areas"]
        if sw["has_sidewalk"]:
            lines += [f"Sidewalk Areas ({sw['location']}):", f"  - Pedestrians: {len(sw['pedestrians'])}", f"  - Static objects: {len(sw['static_objects'])}\n"]
        infra = graph["static_infrastructure"]
        if sum(len(infra[k]) for k in ("barriers", "traffic_cones", "construction", "other")) > 0:
            lines.append("Static Infrastructure:")
            if infra["barriers"]:
                lines.append(f"  - Barriers: {len(infra['barriers'])}")
            if infra["traffic_cones"]:
                lines.append(f"  - Traffic cones: {len(infra['traffic_cones'])}")
            if infra["construction"]:
                lines.append(f"  - Construction: {len(infra['construction'])}\n")
        if graph["spatial_zones"]:
            lines.append("Spatial Zones:")
            lines += [f"  - {z['zone_name']}: {len(z['objects'])} objects (criticality: {z['criticality']})" for z in graph["spatial_zones"] if z["objects"]]
        if graph["safety_critical_elements"]:
            lines.append("\nSafety Critical Elements:")
            lines += [f"  - {e}" for e in graph["safety_critical_elements"]]
        return "\n".join(lines)


def fallback_scene_graph(total_objects: int) -> Dict[str, Any]:
    """`model_dump()` of the minimal graph the reference returns when its LLM call fails (scenegraph_agent.py:381-421)."""
    unknown = "unknown"
    return {"scene_summary": "Error generating scene graph",
            "environment": {"lighting": unknown, "weather": unknown, "visibility_overall": unknown, "location_type": unknown},
            "road_structure": {"road_type": unknown, "lanes": {"lane_count": 0, "lane_type": unknown, "ego_lane_position": unknown, "lane_markings": []},
                               "road_elements": [], "surface_condition": unknown},
            "traffic_participants": {"vehicles": [], "cyclists": [], "vulnerable_road_users": []},
            "sidewalk_areas": {"has_sidewalk": False, "pedestrians": [], "static_objects": [], "location": unknown},
            "static_infrastructure": {"barriers": [], "traffic_cones": [], "construction": [], "other": []},
            "spatial_zones": [], "safety_critical_elements": ["Scene graph generation failed"], "total_objects": int(total_objects)}
