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

_LIDAR_METHODS = ("_preprocess_point_cloud", "_segment_ground", "_generate_multi_layer_bev", "_generate_cluster_visualization")
_SCENE_METHODS = ("_parse_annotations", "_build_spatial_zones")


def patch_reference(lidar_cls=None, scenegraph_cls=None, engine: Optional[GeometryEngine] = None):
    """Monkey-patch the reference classes in place.  Returns the engine the patched methods use."""
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
    if scenegraph_cls is not None:
        scenegraph_cls._table = _wrap(_sg.SceneGraphAgent._table)
        for name in _SCENE_METHODS:
            setattr(scenegraph_cls, name, _wrap(getattr(_sg.SceneGraphAgent, name)))
    return eng
