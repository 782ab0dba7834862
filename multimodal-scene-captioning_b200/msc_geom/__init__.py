"""msc_geom -- B200-native geometric-evidence path for multimodal-scene-captioning.

Host side mirrors the reference's Python interface for this path (loader, LiDARAgent,
SceneGraphAgent, CameraAgent evidence); the work runs in hand-written sm_100a CUDA kernels behind the
C-ABI in include/msc_geom.h.  There is no CPU fallback: importing the engine without the built
extension raises.
"""
__version__ = "0.1.0"
