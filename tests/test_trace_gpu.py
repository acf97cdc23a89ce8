"""Traversal parity: closest-hit records of the CUDA kernels against the CPU restatement of
raytracer.glsl.  north_star tolerance: triangle ids exact except documented ties, t and barycentrics
within 1e-5 relative -- the kernels are in fact bit-identical (same fp32 op order, no FMA), which is
what these tests assert; the 1e-5 bound is asserted too so a future relaxation stays inside it."""
import numpy as np
import pytest

import scenes
from realtimeraytracing_b200 import capi, synth
from realtimeraytracing_b200.layouts import RAY

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5  # north_star


def assert_hits_equal(got, exp, what=""):
    assert np.array_equal(got["did_hit"], exp["did_hit"]), what + " did_hit"
    h = exp["did_hit"] == 1
    assert np.array_equal(got["tri"][h], exp["tri"][h]), what + " triangle ids"
    for f in ("t", "b0", "b1", "b2"):
        a, b = got[f][h], exp[f][h]
        assert np.all(np.abs(a - b) <= REL_TOL * np.maximum(np.abs(b), 1e-30)), what + " " + f
        assert np.array_equal(a, b), what + " %s is not bit-identical" % f
    assert not got["t"][~h].any() and not got["tri"][~h].any()


def build_pair(ctx, oracle, tris, meshes):
    bvh = capi.Bvh(ctx).build(tris, meshes)
    flat = bvh.flat_nodes()
    return bvh, flat


@pytest.mark.parametrize("flags", [capi.TRACE_DEFAULT, capi.TRACE_REFERENCE_ORDER])
def test_primary_rays_soup(ctx, oracle, flags):
    tris, meshes, L = scenes.soup(20000)
    bvh, flat = build_pair(ctx, oracle, tris, meshes)
    try:
        W, H = 160, 96
        cam = synth.soup_camera(L, W, H)
        got = bvh.trace_primary(cam, W, H, W, H, flags=flags)
        exp = oracle.trace_primary(flat, tris, meshes, cam, W, H, W, H)
        assert exp["did_hit"].mean() > 0.3
        assert_hits_equal(got, exp, "soup flags=%d" % flags)
        hits_dev = ctx.dev_alloc(got.nbytes)
        try:  # the asynchronous form reports rays that ran out of stack on request (the host form refuses the frame)
            bvh.trace_primary_dev(cam, W, H, hits_dev, W, H, flags=flags)
            assert bvh.stack_overflows() == 0
        finally:
            ctx.dev_free(hits_dev)
    finally:
        bvh.close()


@pytest.mark.parametrize("flags", [capi.TRACE_DEFAULT, capi.TRACE_REFERENCE_ORDER])
def test_primary_rays_mesh_with_exact_ties(ctx, oracle, flags):
    """Connected mesh: rays through shared edges/vertices give equal t on two leaves (SURVEY Q6 ties)."""
    tris, meshes = synth.grid_mesh(48, 40)
    bvh, flat = build_pair(ctx, oracle, tris, meshes)
    try:
        W, H = 192, 128
        cam = synth.reference_camera(aspect=W / H)
        got = bvh.trace_primary(cam, W, H, W, H, flags=flags)
        exp = oracle.trace_primary(flat, tris, meshes, cam, W, H, W, H)
        assert exp["did_hit"].mean() > 0.4
        assert_hits_equal(got, exp, "mesh flags=%d" % flags)
    finally:
        bvh.close()


def test_reference_pixel_mapping_q5(ctx, oracle):
    """1920x1080-style geometry: floor(H/16)*16 rows are traced and divide the pixel position (Q5)."""
    tris, meshes, L = scenes.soup(5000)
    bvh, flat = build_pair(ctx, oracle, tris, meshes)
    try:
        W, H = 120, 67  # floor -> 112 x 64
        cam = synth.soup_camera(L, W, H)
        got = bvh.trace_primary(cam, W, H)  # denominators default to the reference formula
        dw, dh = synth.reference_denominators(W, H)
        assert (dw, dh) == (112, 64)
        exp = oracle.render(flat, tris, meshes, cam, W, H, dw, dh)[1]
        assert_hits_equal(got, exp, "Q5")
        g = got.reshape(H, W)
        assert not g["did_hit"][dh:, :].any() and not g["did_hit"][:, dw:].any()
    finally:
        bvh.close()


def test_two_meshes_with_model_matrices(ctx, oracle):
    tris, meshes = scenes.two_mesh_scene()
    bvh, flat = build_pair(ctx, oracle, tris, meshes)
    try:
        W, H = 128, 96
        cam = synth.reference_camera(eye=(0.0, 0.0, -9.0), aspect=W / H)
        got = bvh.trace_primary(cam, W, H, W, H)
        exp = oracle.trace_primary(flat, tris, meshes, cam, W, H, W, H)
        assert exp["did_hit"].sum() > 100
        assert_hits_equal(got, exp, "two meshes")
    finally:
        bvh.close()


def test_explicit_ray_batch_and_any_hit(ctx, oracle):
    tris, meshes, L = scenes.soup(8000)
    bvh, flat = build_pair(ctx, oracle, tris, meshes)
    try:
        rng = np.random.RandomState(3)
        n = 5000
        rays = np.zeros(n, dtype=RAY)
        rays["o"][:, :3] = rng.uniform(-L, L, size=(n, 3))
        rays["o"][:, 3] = 1.0
        d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays["d"][:, :3] = d
        rays["d"][::50, 0] = 0.0  # axis-parallel components: +-inf slabs (Q11)
        for flags in (capi.TRACE_DEFAULT, capi.TRACE_REFERENCE_ORDER):
            got = bvh.trace_rays(rays, flags=flags)
            exp = oracle.trace_rays(flat, tris, meshes, rays)
            assert_hits_equal(got, exp, "batch flags=%d" % flags)
        tmax = rng.uniform(0.0, L, size=n).astype(np.float32)
        occ = bvh.trace_rays(rays, any_hit=True, t_max=tmax)
        assert np.array_equal(occ["did_hit"], oracle.any_hit(flat, tris, meshes, rays, tmax))
        occ_inf = bvh.trace_rays(rays, any_hit=True)
        assert np.array_equal(occ_inf["did_hit"], exp["did_hit"])
    finally:
        bvh.close()


@pytest.mark.parametrize("shadow", [False, True])
def test_multi_bounce_render(ctx, oracle, shadow):
    tris, meshes, L = scenes.soup(20000)
    bvh, flat = build_pair(ctx, oracle, tris, meshes)
    try:
        W, H = 96, 64
        cam = synth.soup_camera(L, W, H)
        light = (0.3 * L, 0.8 * L, -1.2 * L)
        for flags in (capi.TRACE_DEFAULT, capi.TRACE_REFERENCE_ORDER):
            rgba, hits, nrays = bvh.render(cam, W, H, W, H, bounces=2, shadow=shadow, light=light, flags=flags)
            ergba, ehits, enrays = oracle.render(flat, tris, meshes, cam, W, H, W, H, bounces=2, shadow=shadow, light=light)
            assert nrays == enrays
            assert_hits_equal(hits, ehits, "render primary")
            assert np.array_equal(rgba, ergba), "image is not bit-identical"
            assert rgba[..., 0].max() > 0.05
    finally:
        bvh.close()


@pytest.mark.parametrize("shadow", [False, True])
def test_shared_walks_of_the_drain_keep_the_records(ctx, oracle, shadow):
    """Once the job pool of the persistent kernel is empty, idle lanes take over half of a walking lane's stack and the
    ray's closest hit is assembled from what its lanes found (trace.cu, drain).  A small frame of long paths on a big
    scene is almost all drain: one job tile per warp, the pool empty from the start.  The frame, the primary records and
    the ray count equal the shader's order (one lane per pixel, no sharing) and, for the smaller scene, the oracle."""
    for n, W, H, bounces, against_oracle in ((20000, 64, 48, 5, True), (400_000, 256, 160, 6, False)):
        tris, meshes, L = scenes.soup(n)
        if against_oracle:
            bvh, flat = build_pair(ctx, oracle, tris, meshes)
        else:
            bvh = capi.Bvh(ctx).build(tris, meshes)
        try:
            cam = synth.soup_camera(L, W, H)
            light = (0.3 * L, 0.8 * L, -1.2 * L)
            a = bvh.render(cam, W, H, W, H, bounces=bounces, shadow=shadow, light=light, flags=capi.TRACE_DEFAULT)
            b = bvh.render(cam, W, H, W, H, bounces=bounces, shadow=shadow, light=light, flags=capi.TRACE_REFERENCE_ORDER)
            assert a[2] == b[2] and a[2] > 1.3 * W * H                    # rays traced: paths do bounce
            assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)), "frame"
            assert np.array_equal(a[1].view(np.uint8), b[1].view(np.uint8)), "primary records"
            if against_oracle:
                ergba, ehits, enrays = oracle.render(flat, tris, meshes, cam, W, H, W, H, bounces=bounces, shadow=shadow, light=light)
                assert a[2] == enrays and np.array_equal(a[0], ergba)
                assert_hits_equal(a[1], ehits, "drain primary")
            # explicit closest-hit rays (the other job kind), a batch small enough to be all drain
            rng = np.random.RandomState(11)
            m = 3000
            rays = np.zeros(m, dtype=RAY)
            rays["o"][:, :3] = rng.uniform(-L, L, size=(m, 3)); rays["o"][:, 3] = 1.0
            d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
            rays["d"][:, :3] = d
            x = bvh.trace_rays(rays, flags=capi.TRACE_DEFAULT)
            y = bvh.trace_rays(rays, flags=capi.TRACE_REFERENCE_ORDER)
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8)) and x["did_hit"].sum() > 100
        finally:
            bvh.close()


def test_row_ranges_tile_the_frame(ctx):
    """Rows [row0,row1) rendered separately equal the full frame (what the multi-GPU sharding relies on)."""
    tris, meshes, L = scenes.soup(10000)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    try:
        W, H = 128, 80
        cam = synth.soup_camera(L, W, H)
        full, fh, fr = bvh.render(cam, W, H, W, H, bounces=1)
        parts, total = [], 0
        for r0 in range(0, H, 16):
            rgba, _, nr = bvh.render(cam, W, H, W, H, row0=r0, row1=min(H, r0 + 16), bounces=1)
            parts.append(rgba); total += nr
        assert np.array_equal(np.concatenate(parts, axis=0), full) and total == fr
    finally:
        bvh.close()


def test_pruned_equals_reference_order_on_a_big_frame(ctx):
    """Size-independent property at scale: the default (pruned) order returns exactly the records of the
    shader's order on a 1M-triangle scene at 1280x720 (no oracle needed)."""
    tris, meshes, L = scenes.soup(1_000_000)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    try:
        W, H = 1280, 720
        cam = synth.soup_camera(L, W, H)
        a = bvh.trace_primary(cam, W, H, flags=capi.TRACE_DEFAULT)
        b = bvh.trace_primary(cam, W, H, flags=capi.TRACE_REFERENCE_ORDER)
        assert a["did_hit"].mean() > 0.5
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    finally:
        bvh.close()


def test_adopted_bvh_traces_identically(ctx):
    tris, meshes, L = scenes.soup(6000)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    try:
        W, H = 64, 48
        cam = synth.soup_camera(L, W, H)
        exp = bvh.trace_primary(cam, W, H, W, H)
        other = capi.Bvh(ctx).adopt_dev(bvh.device_nodes, tris.size, bvh.device_triangles, bvh.device_meshes, meshes.size)
        got = other.trace_primary(cam, W, H, W, H)
        assert np.array_equal(got.view(np.uint8), exp.view(np.uint8))
        other.close()
    finally:
        bvh.close()


@pytest.mark.parametrize("layout", [[1, 1, 1], [3, 8, 8], [0, 4, 4], [5, 8]])
def test_weighted_stripes_tile_the_frame(ctx, layout):
    """Row blocks dealt with weights (rtr_render_stripes_dev): the stripes of all ranks, rendered one after
    the other into one image on this GPU, give the plain frame -- including a rank that owns nothing."""
    tris, meshes, L = scenes.soup(10000)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    try:
        W, H, rpb = 128, 200, 8
        cam = synth.soup_camera(L, W, H)
        full, _, fr = bvh.render(cam, W, H, W, H, bounces=1)
        d_rgba = ctx.dev_alloc(W * H * 16); d_rays = ctx.dev_alloc(8)
        ctx.zero(d_rgba, W * H * 16); ctx.zero(d_rays, 8)
        for rank in range(len(layout)):
            bvh.render_stripes_dev(cam, W, H, d_rgba, rpb, layout, rank, rays_dev=d_rays, denom_w=W, denom_h=H, bounces=1)
        got = np.zeros((H, W, 4), np.float32); rays = np.zeros(1, np.uint64)
        ctx.download(got, d_rgba); ctx.download(rays, d_rays)
        ctx.dev_free(d_rgba); ctx.dev_free(d_rays)
        assert np.array_equal(got, full.reshape(H, W, 4)) and int(rays[0]) == fr
    finally:
        bvh.close()


def test_axis_parallel_rays_take_the_exact_slab_test(ctx, oracle):
    """Rays with a zero direction component: the compressed traversal records cannot bound 1/0, the kernel
    falls back to the reference slab test on the decoded boxes (Q11) -- same records as the oracle."""
    tris, meshes, L = scenes.soup(20000)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    try:
        flat = bvh.flat_nodes()
        rng = np.random.default_rng(5)
        n = 4096
        rays = np.zeros(n, RAY)
        o = rng.uniform(-0.5 * L, 0.5 * L, (n, 3)).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        d[np.arange(n), rng.integers(0, 3, n)] = 0.0          # one component exactly zero
        d[: n // 4, 1] = 0.0                                   # a quarter with two
        d[(d == 0).all(axis=1)] = (1.0, 0.0, 0.0)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        o[:, 2] -= L                                           # from outside, towards the soup
        d[:, 2] = np.abs(d[:, 2])
        rays["o"][:, :3] = o; rays["o"][:, 3] = 1.0
        rays["d"][:, :3] = d
        got = bvh.trace_rays(rays)
        exp = oracle.trace_rays(flat, tris, meshes, rays)
        assert_hits_equal(got, exp, "axis-parallel")
        ref = bvh.trace_rays(rays, flags=capi.TRACE_REFERENCE_ORDER)
        assert np.array_equal(got.view(np.uint8), ref.view(np.uint8))
    finally:
        bvh.close()


@pytest.mark.parametrize("wireframe", [False, True])
def test_shaded_frame_equals_getcolor(ctx, oracle, wireframe):
    """The reference's actual output: primary hits -> getColor (raytracer.glsl:159-179), two models with
    different materials, with and without the wireframe mode."""
    from realtimeraytracing_b200.layouts import MESH
    tris, meshes = synth.grid_mesh(24, 20)
    tris = tris.copy(); tris["model_id"][tris.size // 2:] = 1
    meshes = np.concatenate([meshes, meshes]).view(MESH).copy()
    meshes["material_id"] = [1, 0]
    materials = np.array([[0.2, 0.4, 0.6, 1.0], [0.9, 0.1, 0.3, 0.5]], np.float32)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    try:
        W, H = 96, 64
        cam = synth.reference_camera(aspect=W / H)
        hits = bvh.trace_primary(cam, W, H, W, H)
        assert 0.2 < hits["did_hit"].mean() < 1.0
        got = ctx.shade(hits, tris, meshes, materials, wireframe=wireframe)
        exp = oracle.shade(hits, tris, meshes, materials, wireframe=wireframe)
        assert np.array_equal(got, exp)
        assert len(np.unique(got.view([("c", "<f4", 4)]))) >= (2 if wireframe else 3)
        with pytest.raises(capi.RtrError):
            ctx.shade(hits, tris, meshes, materials[:1], wireframe=wireframe)   # model 0 names material 1
    finally:
        bvh.close()


@pytest.mark.parametrize("depth", [0, 3, 7, 12])
def test_bvh_depth_overlay_equals_the_shaders_loop(ctx, oracle, depth):
    """uIsBVHDisplayed / uDepthDisplayBVH (raytracer.glsl:269-275, edge code 2 of intersectBVH :222-233): the kernel
    finds the left-most intersected node of that depth, the oracle runs the shader's loop to the letter (every
    intersected node of the depth overwrites the colour, the last one visited stays) -- same pixels; then getColor."""
    tris, meshes, L = scenes.soup(6000, seed=3)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    try:
        W, H = 96, 64
        cam = synth.soup_camera(L, W, H)
        flat = bvh.flat_nodes()
        got = bvh.depth_overlay(cam, W, H, depth, W, H)
        exp = oracle.depth_overlay(flat, cam, W, H, depth)
        assert np.array_equal(got, exp)
        colours = {tuple(c) for c in np.unique(got, axis=0).tolist()}
        assert len(colours) >= 2                       # box colour and background at least
        hits = bvh.trace_primary(cam, W, H, W, H)
        mats = np.array([[0.3, 0.6, 0.9, 1.0]], np.float32)
        img = ctx.shade(hits, tris, meshes, mats, wireframe=True, bvh_rgba=got)
        assert np.array_equal(img, oracle.shade(hits, tris, meshes, mats, wireframe=True, bvh_rgba=exp))
    finally:
        bvh.close()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 13, 31, 64])
def test_tiny_trees_through_the_wide_step(ctx, oracle, n):
    """Every shape of the four-slot record near the root: a leaf root, leaf children that occupy a slot as themselves,
    one inner and one leaf child, full grandchildren (bvh.cuh, second record half)."""
    tris, meshes, L = scenes.soup(n, seed=77 + n)
    bvh, flat = build_pair(ctx, oracle, tris, meshes)
    try:
        W, H = 96, 64
        cam = synth.soup_camera(L, W, H)
        got = bvh.trace_primary(cam, W, H, W, H)
        exp = oracle.trace_primary(flat, tris, meshes, cam, W, H, W, H)
        assert_hits_equal(got, exp, "n=%d" % n)
        # incoherent rays from inside and outside the scene, closest hit and any hit with a distance limit
        rng = np.random.default_rng(n)
        rays = np.zeros(4096, dtype=RAY)
        rays["o"][:, :3] = rng.uniform(-1.5 * L, 1.5 * L, (rays.size, 3)).astype(np.float32)
        rays["o"][:, 3] = 1.0
        d = rng.normal(size=(rays.size, 3))
        rays["d"][:, :3] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        got = bvh.trace_rays(rays)
        exp = oracle.trace_rays(flat, tris, meshes, rays)
        assert_hits_equal(got, exp, "n=%d random rays" % n)
        tmax = rng.uniform(-0.2 * L, 2.0 * L, size=rays.size).astype(np.float32)  # some limits are negative: nothing can hit
        occ = bvh.trace_rays(rays, any_hit=True, t_max=tmax)
        assert np.array_equal(occ["did_hit"], oracle.any_hit(flat, tris, meshes, rays, tmax))
    finally:
        bvh.close()


def test_long_directions_take_the_shaders_order(ctx, oracle):
    """The pruning margin is derived for the shader's unit directions (tests/test_prune_bound_cpu.py): a batch of
    rtr_trace_rays that holds a direction longer than 8 is traced in the by-the-letter order instead (any length, same
    records, no pruning)."""
    tris, meshes, L = scenes.soup(2000)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    try:
        flat = bvh.flat_nodes()
        rng = np.random.default_rng(11)
        rays = np.zeros(256, dtype=RAY)
        rays["o"][:, :3] = rng.uniform(-0.4 * L, 0.4 * L, (rays.size, 3)).astype(np.float32)
        rays["o"][:, 2] = -2.0 * L
        rays["o"][:, 3] = 1.0
        rays["d"][:, :3] = (rng.normal(size=(rays.size, 3)) * 0.05 + (0.0, 0.0, 1.0)).astype(np.float32)
        rays["d"][5, :3] *= 50.0                                   # |d| ~ 50
        got = bvh.trace_rays(rays)
        exp = oracle.trace_rays(flat, tris, meshes, rays)
        assert exp["did_hit"].sum() > 20
        assert_hits_equal(got, exp, "long direction in the batch")
        rays["d"][5, :3] /= 50.0
        assert_hits_equal(bvh.trace_rays(rays), oracle.trace_rays(flat, tris, meshes, rays), "unit-ish directions")
    finally:
        bvh.close()


def _chain_scene(n):
    """n small triangles facing -z, strung along +z, under a RIGHT-deep chain in DFS pre-order: inner node i sits at
    2i, its left child (the leaf of triangle i, at 2i + 1) is the FARTHEST remaining triangle and its right child
    (2i + 2) holds all the nearer ones -- so both visiting orders stack one entry per level."""
    from realtimeraytracing_b200.layouts import NODE, TRIANGLE
    tris = np.zeros(n, dtype=TRIANGLE)
    z = (n - 1 - np.arange(n)).astype(np.float32)                 # triangle i at z = n - 1 - i
    tris["p0"][:, :3] = np.stack([np.full(n, -1.0), np.full(n, -1.0), z], 1)
    tris["p1"][:, :3] = np.stack([np.full(n, 0.0), np.full(n, 1.0), z], 1)   # host P1, P2 counter-clockwise seen from -z (Q8)
    tris["p2"][:, :3] = np.stack([np.full(n, 1.0), np.full(n, -1.0), z], 1)
    for k in ("p0", "p1", "p2"):
        tris[k][:, 3] = 1.0
    flat = np.zeros(2 * n - 1, dtype=NODE)
    for i in range(n - 1):
        flat[2 * i]["bmin"] = (-1.0, -1.0, 0.0)
        flat[2 * i]["bmax"] = (1.0, 1.0, z[i])                    # triangles i .. n-1 lie at z <= z[i]
        flat[2 * i]["left"], flat[2 * i]["right"] = 2 * i + 1, 2 * i + 2
    for i in range(n):
        p = 2 * i + 1 if i < n - 1 else 2 * n - 2
        flat[p]["bmin"] = (-1.0, -1.0, z[i])
        flat[p]["bmax"] = (1.0, 1.0, z[i])
        flat[p]["tri"] = i
    return tris, synth.identity_meshes(1), flat


def test_stack_overflow_is_never_silent(ctx, oracle):
    """Chains deeper than the 128-entry lane stack.  The shader's own stack holds 1024 entries (raytracer.glsl:251):
    up to that depth the host-pointer calls return the oracle's records -- they fall back to RTR_TRACE_DEEP_STACK, the
    shader's loop with the shader's stack -- and beyond it (where the shader itself writes out of bounds) they refuse;
    explicit ray batches and any-hit rays included.  The asynchronous forms report the count."""
    rays = np.zeros(64, dtype=RAY)
    rays["o"][:, :3] = (0.05, -0.1, -10.0); rays["o"][:, 3] = 1.0
    rays["d"][:, :3] = (0.001, 0.002, 1.0)
    rays["d"][:, :3] /= np.linalg.norm(rays["d"][0, :3])
    # through a corner of every box but outside every triangle: an any-hit ray finds no occluder and has to walk the
    # whole chain (with a distance limit the default order would prune the chain away instead)
    miss = rays.copy()
    miss["o"][:, :3] = (0.9, 0.9, -10.0)
    miss["d"][:, :3] = (0.0001, 0.0001, 1.0)
    miss["d"][:, :3] /= np.linalg.norm(miss["d"][0, :3])
    for depth, outcome in ((100, "fits"), (300, "deep stack"), (1300, "refused")):
        tris, meshes, flat = _chain_scene(depth)
        d_nodes, d_tris, d_meshes = ctx.dev_alloc(flat.nbytes), ctx.dev_alloc(tris.nbytes), ctx.dev_alloc(meshes.nbytes)
        ctx.upload(d_nodes, flat); ctx.upload(d_tris, tris); ctx.upload(d_meshes, meshes)
        bvh = capi.Bvh(ctx).adopt_dev(d_nodes, tris.size, d_tris, d_meshes, meshes.size)
        try:
            if outcome != "refused":   # the oracle's stack is the shader's: 1024 entries
                exp = oracle.trace_rays(flat, tris, meshes, rays)
                assert exp["did_hit"].all() and (exp["tri"] == tris.size - 1).all()        # the nearest triangle, z = 0
                assert not oracle.trace_rays(flat, tris, meshes, miss)["did_hit"].any()
            for flags in (capi.TRACE_DEFAULT, capi.TRACE_REFERENCE_ORDER, capi.TRACE_DEEP_STACK):
                for any_hit in (False, True):
                    batch = miss if any_hit else rays
                    if outcome == "refused":
                        with pytest.raises(capi.RtrError) as e:
                            bvh.trace_rays(batch, any_hit=any_hit, flags=flags)
                        assert e.value.code == -5 and "1024" in str(e.value), (flags, any_hit, str(e.value))   # RTR_E_UNSUPPORTED
                    else:
                        got = bvh.trace_rays(batch, any_hit=any_hit, flags=flags)
                        if any_hit: assert not got["did_hit"].any()
                        else: assert_hits_equal(got, exp, "chain of %d, flags %d" % (depth, flags))
            d_rays, d_hits = ctx.dev_alloc(rays.nbytes), ctx.dev_alloc(rays.size * 24)
            ctx.upload(d_rays, rays)
            bvh.trace_rays_dev(d_rays, rays.size, d_hits)
            assert bvh.stack_overflows() == (0 if outcome == "fits" else rays.size)
            bvh.trace_rays_dev(d_rays, rays.size, d_hits, flags=capi.TRACE_DEEP_STACK)
            assert bvh.stack_overflows() == (rays.size if outcome == "refused" else 0)
            ctx.dev_free(d_rays); ctx.dev_free(d_hits)
        finally:
            bvh.close()
            ctx.dev_free(d_nodes); ctx.dev_free(d_tris); ctx.dev_free(d_meshes)


def test_malformed_adopted_arrays_are_refused_without_a_device_fault(ctx, oracle):
    """rtr_bvh_adopt_dev on a foreign node array: links that leave the array or triangle ids that do not exist are
    counted and never followed -- RTR_E_UNSUPPORTED, and the context stays usable (no sticky CUDA error)."""
    tris, meshes, L = scenes.soup(500)
    good = capi.Bvh(ctx).build(tris, meshes)
    flat = good.flat_nodes()
    good.close()
    d_tris, d_meshes = ctx.dev_alloc(tris.nbytes), ctx.dev_alloc(meshes.nbytes)
    ctx.upload(d_tris, tris); ctx.upload(d_meshes, meshes)
    try:
        for what in ("right link beyond the array", "left link beyond the array", "triangle id beyond the array", "right link backwards"):
            bad = flat.copy()
            inner = np.nonzero((bad["left"] != 0) | (bad["right"] != 0))[0]
            leaves = np.nonzero((bad["left"] == 0) & (bad["right"] == 0))[0]
            if what.startswith("right link beyond"):
                bad["right"][inner[7]] = 0x7FFFFFF0
            elif what.startswith("left"):
                bad["left"][inner[3]] = 2 * tris.size + 5
            elif what.startswith("triangle"):
                bad["tri"][leaves[11]] = 0xFFFFFF00
            else:
                bad["right"][inner[9]] = 1
            d_nodes = ctx.dev_alloc(bad.nbytes)
            ctx.upload(d_nodes, bad)
            with pytest.raises(capi.RtrError) as e:
                capi.Bvh(ctx).adopt_dev(d_nodes, tris.size, d_tris, d_meshes, meshes.size)
            assert e.value.code == -5, (what, str(e.value))
            ctx.dev_free(d_nodes)
        # the context is still healthy: the unmodified array adopts and traces like the build it came from
        d_nodes = ctx.dev_alloc(flat.nbytes)
        ctx.upload(d_nodes, flat)
        bvh = capi.Bvh(ctx).adopt_dev(d_nodes, tris.size, d_tris, d_meshes, meshes.size)
        W, H = 64, 48
        cam = synth.soup_camera(L, W, H)
        assert_hits_equal(bvh.trace_primary(cam, W, H, W, H), oracle.trace_primary(flat, tris, meshes, cam, W, H, W, H), "adopted")
        bvh.close()
        ctx.dev_free(d_nodes)
    finally:
        ctx.dev_free(d_tris); ctx.dev_free(d_meshes)
