"""CPU oracle for the acceleration-structure + ray-cast path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
``realtimeraytracing_b200`` never does (tests/test_no_oracle_in_product.py enforces it).

* ``Oracle``    -- ctypes binding of oracle/librtr_oracle.so (rtr_oracle.c, the C restatement)
* ``Reference`` -- ctypes binding of oracle/_ref/libref_bvh_<cap>.so (the reference's own
                   bvh.cpp/triangle.cpp compiled by oracle/Makefile; build half)
* ``ReferenceRaytracer`` -- ctypes binding of oracle/_ref/libref_raytracer_<variant>.so (the reference's own
                   raytracer.glsl compiled as C++ through its vendored GLM; traversal half)
"""
from .bindings import (Oracle, Reference, ReferenceRaytracer, build_oracle, build_reference,  # noqa: F401
                       oracle_available, raytracer_available, reference_available)
