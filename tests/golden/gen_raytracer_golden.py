"""Generates tests/golden/raytracer_golden.npz from the REFERENCE's own raytracer.glsl, compiled as C++ through the
reference's vendored GLM (oracle/ref_raytracer.cpp, oracle/glsl2cpp.sed -> oracle/_ref/libref_raytracer_*.so).
Run in the build container only (needs /root/reference):

    python tests/golden/gen_raytracer_golden.py

Scenes, cameras and ray batches are regenerated from seeds by tests/raytracer_cases.py and the BVH by the (pinned)
build half of the oracle, so the fixture holds shader OUTPUTS only, per case:
  rays_<v>      primary rays of main() (:303-305 + getRay :92-100), direction words          v in {div, glm}
  hits_<v>      getClosestHitBVH (:246-295) of those rays, 6 words per ray
  allhits_div   getAllHits (:149-157) of the same rays (every 7th ray: it is brute force)
  image_div     main() through glDispatchCompute(floor(W/16), floor(H/16)) with the BVH overlay at depth 3 and
                wireframe on; NaN where no invocation stores (Q5)
  batch_hits_div   getClosestHitBVH of a seeded random ray batch
  batch_tri_div    rayTriangleIntersection (:102-147) of rays aimed at / around / grazing seeded triangles
  batch_box_div    intersectBVH (:182-237, codes 0/1/2 with uDepthDisplayBVH = 3) of rays aimed around seeded nodes
"div" pins normalize := v / sqrt(dot(v, v)) (what rtr_oracle.c and the kernels use): compared bit for bit.
"glm" is glm::normalize = v * (1 / sqrt(dot)): compared within the 1e-5 relative tolerance of BASELINE.json.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import raytracer_cases as rc  # noqa: E402
from oracle import Oracle, ReferenceRaytracer  # noqa: E402

OVERLAY_DEPTH = 3
BATCH = 4000


def main():
    o = Oracle()
    out = {}
    for name in rc.CASE_NAMES:
        tris, meshes, materials, cam, w, h = rc.case(name)
        b = o.bvh_build(tris, meshes)
        flat = o.flatten(b.clusters, b.left, b.right)
        dw, dh = (w // 16) * 16, (h // 16) * 16
        for v in ("div", "glm"):
            r = ReferenceRaytracer(v)
            r.set_scene(tris, meshes, flat, materials)
            r.set_camera(cam)
            r.set_flags(-1, False, False)
            rays = r.primary_rays(w, h, dw, dh)
            out["%s/rays_%s" % (name, v)] = rays["d"].view(np.uint32).copy()
            out["%s/hits_%s" % (name, v)] = r.closest_hit_bvh(rays).view(np.uint32).reshape(-1, 6).copy()
            if v == "div":
                out[name + "/allhits_div"] = r.all_hits(rays[::7]).view(np.uint32).reshape(-1, 6).copy()
                r.set_flags(OVERLAY_DEPTH, True, True)
                out[name + "/image_div"] = r.dispatch(w, h)
                r.set_flags(-1, False, False)
                extent = float(np.abs(np.concatenate([tris["p0"][:, :3], tris["p1"][:, :3], tris["p2"][:, :3]])).max()) * 2
                batch = rc.random_rays(BATCH, extent, seed=len(name))
                rng = np.random.RandomState(5)
                ti = rng.randint(0, tris.size, size=BATCH).astype(np.uint32)
                ni = rng.randint(0, flat.size, size=BATCH).astype(np.uint32)
                out[name + "/batch_hits_div"] = r.closest_hit_bvh(batch).view(np.uint32).reshape(-1, 6).copy()
                out[name + "/batch_tri_div"] = r.ray_triangle(rc.rays_at_triangles(tris, meshes, ti, 6), ti) \
                    .view(np.uint32).reshape(-1, 6).copy()
                r.set_flags(OVERLAY_DEPTH, False, False)   # the edge code 2 depends on uDepthDisplayBVH (:223)
                out[name + "/batch_box_div"] = r.intersect_bvh(rc.rays_at_boxes(flat, ni, 7), ni).astype(np.uint8)
        hits = out[name + "/hits_div"]
        print(name, tris.size, "hit rate %.3f" % (hits[:, 4] != 0).mean(),
              "batch hit rate %.3f" % (out[name + "/batch_hits_div"][:, 4] != 0).mean(),
              "tri hits %d" % int((out[name + "/batch_tri_div"][:, 4] != 0).sum()),
              "box codes", np.bincount(out[name + "/batch_box_div"], minlength=3))
    np.savez_compressed(os.path.join(HERE, "raytracer_golden.npz"), **out)


if __name__ == "__main__":
    main()
