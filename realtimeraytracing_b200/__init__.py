"""realtimeraytracing_b200 -- B200-native (sm_100a) acceleration-structure + ray-cast path of
MrBigoudi/RealTimeRaytracing behind a C ABI (include/rtr.h).

Host-side mirror of the reference interface for this path:

* ``capi.Context`` / ``capi.Bvh``  -- thin ctypes binding of librtr_b200.so
* ``scene.BVH``                    -- cr::BVH-shaped object (ctor builds, ``_InternalStruct`` holds
                                      BVH_Params) and ``scene.flatten`` (glr::Scene::getBVH_NodesToGPUData)
* ``parallel``                     -- one process per GPU: BVH broadcast + image-row sharding
* ``synth`` / ``layouts``          -- deterministic inputs and the reference's record layouts

There is no CPU implementation in this package: every entry point runs CUDA kernels or raises.
"""
from . import layouts, synth  # noqa: F401
from .capi import Bvh, Context, RtrError, TRACE_DEFAULT, TRACE_REFERENCE_ORDER, load_library  # noqa: F401

__version__ = "0.1.0"
