#!/bin/bash
# compute-sanitizer over a small end-to-end pass (sort, rebuild, flatten, default + reference-order traversal, shading)
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build > gpurun_out/build.log 2>&1
cat > /tmp/san_driver.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from realtimeraytracing_b200 import capi, synth
with capi.Context(0) as ctx:
    keys = synth.random_keys_u32(40000, seed=3)
    vals = np.arange(keys.size, dtype=np.uint32)
    k2, v2 = ctx.sort_pairs_u32(keys, vals)
    assert np.array_equal(k2, np.sort(keys, kind="stable"))
    assert np.array_equal(ctx.sort_keys_u32(keys[:7000]), np.sort(keys[:7000]))
    tris, meshes, L = synth.triangle_soup(30000)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    W, H = 96, 64
    cam = synth.soup_camera(L, W, H)
    a = bvh.trace_primary(cam, W, H, W, H)
    b = bvh.trace_primary(cam, W, H, W, H, flags=capi.TRACE_REFERENCE_ORDER)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    rgba, hits, nr = bvh.render(cam, W, H, W, H, bounces=2, shadow=True, light=(0.0, 2 * L, -2 * L))
    mats = np.array([[1, 1, 1, 1]], np.float32)
    img = ctx.shade(a, tris, meshes, mats, wireframe=True)
    ov = bvh.depth_overlay(cam, W, H, 5)
    img = ctx.shade(a, tris, meshes, mats, bvh_rgba=ov)
    bvh.close()
    for n in (1, 2, 3, 5, 9):  # every slot shape of the wide record near the root
        t2, m2, L2 = synth.triangle_soup(n)
        small = capi.Bvh(ctx).build(t2, m2)
        small.trace_primary(synth.soup_camera(L2, 32, 32), 32, 32, 32, 32)
        small.close()
print("driver ok")
PY
for tool in ${TOOLS:-memcheck racecheck}; do
  timeout ${SAN_TIMEOUT:-1200} compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_driver.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|driver ok|Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
