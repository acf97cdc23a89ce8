"""The facing test of ray_triangle_w (trace.cu) skips the shader's normalisation -- one square root and three divisions
per leaf test -- when |dot3(n, d)|^2 > 4.1e-12 |n|^2 |d|^2: then the plain product must have the sign of the shader's
dot(normalize(n), d) (raytracer.glsl:112-119).  Checked here on the model of the two expressions in numpy float32
(one rounded IEEE op per operation, the kernel's operation order: dot3 = (x*x' + y*y') + z*z', normalize = v / sqrt(dot3))
over adversarial pairs: n . d within a few ulps of zero, at all magnitudes the guard lets through."""
import numpy as np

F = np.float32


def dot3(ax, ay, az, bx, by, bz):
    return (ax * bx + ay * by) + az * bz


def shader_rejects(nx, ny, nz, dx, dy, dz):
    with np.errstate(all="ignore"):
        length = np.sqrt(dot3(nx, ny, nz, nx, ny, nz))
        return dot3(nx / length, ny / length, nz / length, dx, dy, dz) >= F(0)


def test_fast_facing_test_has_the_shaders_sign():
    rng = np.random.default_rng(31)
    m = 400000
    total_fast = 0
    for scale_n, scale_d in [(1, 1), (1e-6, 1), (1e6, 1), (1, 1e-6), (1, 1e6), (1e-12, 1e-3), (1e12, 1e3), (3e-15, 1), (1, 3e14)]:
        n = rng.normal(size=(m, 3))
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        # d = a vector orthogonal to n + a tiny component along n: n . d is 0 ... 1e-4 |n||d| in either direction
        t = rng.normal(size=(m, 3))
        t -= n * np.sum(t * n, axis=1, keepdims=True)
        t /= np.linalg.norm(t, axis=1, keepdims=True)
        eps = rng.choice([-1.0, 1.0], size=(m, 1)) * 10.0 ** rng.uniform(-9, -4, size=(m, 1))
        eps[: m // 20] = 0.0
        d = t + eps * n
        nx, ny, nz = (np.asarray(n[:, k] * scale_n, dtype=F) for k in range(3))
        dx, dy, dz = (np.asarray(d[:, k] * scale_d, dtype=F) for k in range(3))
        with np.errstate(all="ignore"):
            sd = dot3(nx, ny, nz, dx, dy, dz)
            n2 = dot3(nx, ny, nz, nx, ny, nz)
            d2 = dot3(dx, dy, dz, dx, dy, dz)
            bound = F(4.1e-12) * n2 * d2
            fast = (n2 > F(1e-30)) & (n2 < F(1e30)) & (d2 > F(1e-30)) & (d2 < F(1e30)) & (sd * sd > bound)
        slow = shader_rejects(nx, ny, nz, dx, dy, dz)
        assert np.array_equal((sd > F(0))[fast], slow[fast]), (scale_n, scale_d)
        total_fast += int(fast.sum())
        # and the exact sign agrees too, where float64 can tell
        exact = np.sum(np.stack([nx, ny, nz], 1).astype(np.float64) * np.stack([dx, dy, dz], 1).astype(np.float64), axis=1)
        assert np.array_equal((sd > F(0))[fast], (exact > 0)[fast])
    assert total_fast > 1000000  # the guard lets most of these near-perpendicular pairs through
