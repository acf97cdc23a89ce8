"""The default traversal order skips a box whose entry distance exceeds (t_best + K*E)(1 + 2K), K = 64 u E^2 / 1e-4,
E = longest triangle edge (trace.cu: make_prune_bound).  That is safe iff the computed hit distance of any triangle
the shader accepts (|a| >= 1e-4, raytracer.glsl:117) is within K (t_true + E) of the true one.  Measured here on the
float32 model of rayTriangleIntersection (the kernel's operation order, one rounded op per operation) against float64,
on nearly degenerate triangles just above the acceptance threshold: the error stays below a tenth of the bound for
unit directions and grows with |d| -- which is why rtr_trace_rays refuses |d| > 8 in the default order."""
import numpy as np

F = np.float32
U = 2.0 ** -24


def cross(ax, ay, az, bx, by, bz):
    return ay * bz - by * az, az * bx - bz * ax, ax * by - bx * ay


def dot3(ax, ay, az, bx, by, bz):
    return (ax * bx + ay * by) + az * bz


def ray_triangle(o, d, p0, p1, p2, dt):
    """raytracer.glsl:102-147 in shader naming, without the facing test (it does not change t)."""
    o, d, p0, p1, p2 = (x.astype(dt) for x in (o, d, p0, p1, p2))
    e0 = [p1[:, k] - p0[:, k] for k in range(3)]
    e1 = [p2[:, k] - p0[:, k] for k in range(3)]
    q = cross(d[:, 0], d[:, 1], d[:, 2], *e1)
    a = dot3(*e0, *q)
    with np.errstate(all="ignore"):
        s = [(o[:, k] - p0[:, k]) / a for k in range(3)]
        r = cross(*s, *e0)
        bx = dot3(*s, *q)
        by = dot3(*r, d[:, 0], d[:, 1], d[:, 2])
        bz = (dt(1) - bx) - by
        t = dot3(*e1, *r)
    return (np.abs(a) >= dt(1e-4)) & (bx >= 0) & (by >= 0) & (bz >= 0) & (t >= 0), t, a


def worst_ratio(rng, dlen, E, m=400000):
    p0 = rng.uniform(-1, 1, (m, 3)) * 5
    e0 = rng.normal(size=(m, 3))
    e0 *= rng.uniform(0.2, 1, (m, 1)) * E / np.linalg.norm(e0, axis=1, keepdims=True)
    e1 = e0 * rng.uniform(0.1, 1, (m, 1)) + rng.normal(size=(m, 3)) * E * 10.0 ** rng.uniform(-4, 0, (m, 1))
    n1 = np.linalg.norm(e1, axis=1, keepdims=True)
    e1 = np.where(n1 > E, e1 / n1 * E, e1)
    p1, p2 = p0 + e0, p0 + e1
    w = rng.dirichlet([1, 1, 1], m)
    target = p0 * w[:, :1] + p1 * w[:, 1:2] + p2 * w[:, 2:]
    dirs = rng.normal(size=(m, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    o = target - dirs * 10.0 ** rng.uniform(-3, 1.7, (m, 1))
    args32 = [x.astype(F) for x in (o, dirs * dlen, p0, p1, p2)]
    ok32, t32, a32 = ray_triangle(*args32, F)
    _, t64, _ = ray_triangle(*[x.astype(np.float64) for x in args32], np.float64)
    edges = np.maximum(np.linalg.norm((args32[3] - args32[2]).astype(np.float64), axis=1),
                       np.linalg.norm((args32[4] - args32[2]).astype(np.float64), axis=1))
    emax = edges.max()
    k = 64 * U * emax ** 2 / 1e-4
    sel = ok32 & np.isfinite(t32)
    assert sel.sum() > 5000 and (np.abs(a32[sel]) < 3e-4).sum() > 1000  # plenty of triangles right above the threshold
    return float((np.abs(t32[sel].astype(np.float64) - t64[sel]) / (k * (np.abs(t64[sel]) + emax))).max())


def test_hit_distance_error_stays_inside_the_pruning_bound():
    rng = np.random.default_rng(1)
    for E in (0.05, 0.3, 1.0):
        assert worst_ratio(rng, 1.0, E) < 0.1     # the shader's rays: unit directions
        assert worst_ratio(rng, 0.1, E) < 0.1     # shorter directions are no worse
        assert worst_ratio(rng, 8.0, E) < 0.5     # the longest direction rtr_trace_rays lets through


def test_the_error_grows_with_the_direction_length():
    rng = np.random.default_rng(2)
    assert worst_ratio(rng, 100.0, 0.05) > 1.0    # why longer directions are refused in the default order
