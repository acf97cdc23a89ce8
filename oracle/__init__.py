"""CPU oracle for the acceleration-structure + ray-cast path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
``realtimeraytracing_b200`` never does (tests/test_no_oracle_in_product.py enforces it).

* ``Oracle``    -- ctypes binding of oracle/librtr_oracle.so (rtr_oracle.c, the C restatement)
* ``Reference`` -- ctypes binding of oracle/_ref/libref_bvh_<cap>.so (the reference's own
                   bvh.cpp/triangle.cpp compiled by oracle/Makefile; build half only)
"""
from .bindings import Oracle, Reference, build_oracle, build_reference, oracle_available, reference_available  # noqa: F401
