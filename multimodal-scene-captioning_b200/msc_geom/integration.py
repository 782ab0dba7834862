"""Drop-in wiring into the reference project: replaces the local-geometry methods of the reference's
own LiDARAgent / SceneGraphAgent classes with the GPU-backed ones, leaving every LLM call and the pipeline untouched.

    from agents.content_transform.lidar_agent import LiDARAgent
    from agents.content_transform.scenegraph_agent import SceneGraphAgent
    import msc_geom.integration as mi
    mi.patch_reference(LiDARAgent, SceneGraphAgent)
"""
from __future__ import annotations

from typing import Optional

from . import lidar_agent as _la
from . import scenegraph_agent as _sg
from .engine import GeometryEngine

# local-geometry methods swapped in; `_classify_batch_with_llm`, `call_llm` and everything else remote stay the reference's own
_LIDAR_METHODS = ("_preprocess_point_cloud", "_segment_ground", "_generate_multi_layer_bev", "_generate_cluster_visualization",
                  "_cluster_visualizations", "_cluster_metadata", "_detect_objects_3d")
_SCENE_METHODS = ("_parse_annotations", "_build_spatial_zones")


def patch_reference(lidar_cls=None, scenegraph_cls=None, engine: Optional[GeometryEngine] = None, camera_cls=None):
    """Monkey-patch the reference classes in place.  Returns the engine the patched methods use.
    LiDARAgent: filter / split / BEV, and `_detect_objects_3d` (device DBSCAN + all cluster views in one launch), which hands each batch
    of ten to the reference's OWN `_classify_batch_with_llm(cluster_images, cluster_metadata)`.  SceneGraphAgent: the annotation table.
    CameraAgent: `process` gains the additive `sample` argument whose box -> camera evidence is merged into `context`."""
    eng = engine or GeometryEngine()

    def _ensure(self):
        if getattr(self, "engine", None) is None:
            self.engine = eng
        if not hasattr(self, "_last_table"):
            self._last_table = None

    def _wrap(fn):
        def method(self, *a, **k):
            _ensure(self)
            return fn(self, *a, **k)
        method.__name__, method.__doc__ = fn.__name__, fn.__doc__
        return method

    if lidar_cls is not None:
        lidar_cls._params = _la.LiDARAgent._params
        for name in _LIDAR_METHODS:
            setattr(lidar_cls, name, _wrap(getattr(_la.LiDARAgent, name)))
    if camera_cls is not None:
        from .camera_agent import evidence_context
        ref_process = camera_cls.process

        def process(self, images, camera_names, context=None, sample=None):
            if sample is not None:
                context = {**(context or {}), **evidence_context(eng, sample)}
            return ref_process(self, images, camera_names, context)
        process.__doc__ = ref_process.__doc__
        camera_cls.process = process
    if scenegraph_cls is not None:
        scenegraph_cls._table = _wrap(_sg.SceneGraphAgent._table)
        for name in _SCENE_METHODS:
            setattr(scenegraph_cls, name, _wrap(getattr(_sg.SceneGraphAgent, name)))
    return eng
