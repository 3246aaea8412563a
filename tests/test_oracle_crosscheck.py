"""Second opinion on the oracle: the reference's primitive tests, sampling warps and BSDF branches
restated once more, independently, in numpy float32 (every numpy float32 operation is one IEEE
rounding, evaluated here in the GLSL source order), and compared bit for bit with what the C++ oracle
computes on the same inputs.  The reference ships no test vectors, so two independent
restatements agreeing is the strongest pin available for its arithmetic."""
import ctypes as C

import numpy as np
import pytest

import _libs
from hijiki_b200 import _abi

F = np.float32
EPS = F(1e-4)


def _dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def _cross(a, b):
    return np.array([a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]], dtype=F)


def _triangle(o, d, tmin, tmax, a, b, c):
    """shapes/triangle.glsl:15-52"""
    ab, ac = b - a, c - a
    n = _cross(ab, ac)
    ro = o - a
    q = _cross(ro, d)
    with np.errstate(all="ignore"):
        inv = F(1.0) / _dot(d, n)
        u = inv * _dot(-q, ac)
        v = inv * _dot(q, ab)
        if u < 0 or v < 0 or u + v > F(1.0):
            return None
        t = inv * _dot(-n, ro)
    if tmin <= t <= tmax:
        return t, u, v
    return None


def _sphere(o, d, tmin, tmax, centre, r):
    """shapes/sphere.glsl:18-41"""
    l = o - centre
    b = F(2.0) * _dot(d, l)
    c = _dot(l, l) - r * r
    disc = b * b - F(4.0) * c
    if disc < 0:
        return None
    disc = np.sqrt(disc)
    t0 = F(-0.5) * (b + disc)
    if tmin <= t0 <= tmax:
        return t0
    t1 = F(-0.5) * (b - disc)
    if tmin <= t1 <= tmax:
        return t1
    return None


def _one_shape_scene(kind, rng):
    cam = ((0.0, 0.0, 5.0), (0.0, 0.0, 0.0, 1.0), 40.0)
    if kind == "triangle":
        verts = np.zeros((3, 8), F)
        verts[:, :3] = rng.standard_normal((3, 3)).astype(F)
        verts[:, 4:7] = (0, 0, 1)
        return _libs.CustomScene(cam, triangles=[(0, 1, 2)], vertices=verts, materials=[(_abi.MAT_DIFFUSE, 0)],
                                 diffuse=[(0.5, 0.5, 0.5, 0)]), verts
    sph = np.array([[*rng.standard_normal(3) * 0.3, 0.2 + rng.random()]], F)
    return _libs.CustomScene(cam, spheres=sph, materials=[(_abi.MAT_DIFFUSE, 0)], diffuse=[(0.5, 0.5, 0.5, 0)]), sph


def _rays_towards(rng, n, spread):
    rays = np.zeros(n, dtype=_abi.RAY_DTYPE)
    o = (rng.standard_normal((n, 3)) * 2.0).astype(F)
    target = (rng.standard_normal((n, 3)) * spread).astype(F)
    d = target - o
    d = (d / np.linalg.norm(d.astype(np.float64), axis=1, keepdims=True)).astype(F)
    rays["origin"], rays["direction"] = o, d
    rays["t_min"] = 2e-4
    rays["t_max"] = np.where(rng.random(n) < 0.5, np.inf, rng.random(n) * 4).astype(F)
    return rays


@pytest.mark.parametrize("kind", ["triangle", "sphere"])
def test_primitive_tests_agree_with_numpy_restatement(oracle, kind):
    rng = np.random.default_rng(77 if kind == "triangle" else 78)
    for _ in range(4):
        scene, geom = _one_shape_scene(kind, rng)
        rays = _rays_towards(rng, 1500, 0.6)
        n = rays.size
        ids, t, uv = np.zeros(n, np.int32), np.zeros(n, F), np.zeros((n, 2), F)
        assert oracle.orc_trace(C.byref(scene.view), _libs.ptr(rays), n, 0, 1e-4, _libs.ptr(ids), _libs.ptr(t),
                                _libs.ptr(uv), None, 0) == 0
        hits = 0
        for i in range(n):
            o, d = rays["origin"][i], rays["direction"][i]
            tmin, tmax = rays["t_min"][i], rays["t_max"][i]
            if kind == "triangle":
                r = _triangle(o, d, tmin, tmax, geom[0, :3], geom[1, :3], geom[2, :3])
                if r is None:
                    assert ids[i] == -1
                else:
                    hits += 1
                    assert ids[i] == 0
                    assert F(r[0]).view(np.uint32) == t[i].view(np.uint32)
                    assert F(r[1]).view(np.uint32) == uv[i, 0].view(np.uint32)
                    assert F(r[2]).view(np.uint32) == uv[i, 1].view(np.uint32)
            else:
                r = _sphere(o, d, tmin, tmax, geom[0, :3], geom[0, 3])
                if r is None:
                    assert ids[i] == -1
                else:
                    hits += 1
                    assert ids[i] == 0 and F(r).view(np.uint32) == t[i].view(np.uint32)
        assert 50 < hits < n - 50


def test_rng_and_float_conversion_agree_with_numpy(oracle):
    """rand.glsl:1-20 in numpy uint32 / float32."""
    rng = np.random.default_rng(5)
    for seed in rng.integers(0, 2 ** 32, 200, dtype=np.uint64):
        s = np.uint32(seed)
        with np.errstate(over="ignore"):
            h = (s ^ np.uint32(61)) ^ (s >> np.uint32(16))
            h = h * np.uint32(9)
            h = h ^ (h >> np.uint32(4))
            h = h * np.uint32(0x27d4eb2d)
            h = h ^ (h >> np.uint32(15))
        assert oracle.orc_seed_rng(int(s)) == int(h)
        st = C.c_uint32(int(h))
        x = np.uint32(h)
        for _ in range(3):
            x = x ^ (x << np.uint32(13))
            x = x ^ (x >> np.uint32(17))
            x = x ^ (x << np.uint32(5))
            f = oracle.orc_rand_uniform_float(C.byref(st))
            assert st.value == int(x)
            assert F(f).view(np.uint32) == (F(x) * F(1.0 / 4294967296.0)).view(np.uint32)


def test_first_bounce_of_a_diffuse_path_agrees_with_numpy(oracle, cbox):
    """One full bounce by hand — camera ray (KAT-checked elsewhere), closest triangle by brute force in
    numpy, interpolated normal and tangent frame (shapes/triangle.glsl:54-78) — against the oracle's
    path log (first vertex id and t) and its first-pass feature layer (normal, depth)."""
    tris = cbox.array("triangles")
    verts = cbox.array("vertices")
    w, h = 24, 18
    blocks = _libs.generate_blocks(_libs.hosttest(), w, h, 1, block_size=64)
    op = _libs.orc_params(max_bounces=1, use_bvh=0, block_size=64)
    layers = np.zeros((3, h, w, 4), F)
    assert oracle.orc_integrate_frame(C.byref(cbox.view), _libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(layers),
                                      None, 0) == 0
    so = blocks[0]["sample_offset"]
    one = np.zeros(1, dtype=_abi.RAY_DTYPE)
    checked = 0
    for (px, py) in [(3, 4), (12, 9), (20, 15), (7, 16), (16, 2)]:
        oracle.orc_camera_ray(cbox.view.scene.ptr, float(F(px) + so[0]), float(F(py) + so[1]), float(w), float(h), 1e-4,
                              _libs.ptr(one))
        o, d = one[0]["origin"], one[0]["direction"]
        best = None
        tmax = F(np.inf)
        for ti in range(tris.shape[0]):  # linear scan, scene.glsl:150-156
            a, b, c = (verts[tris[ti, k], :3] for k in range(3))
            r = _triangle(o, d, F(1e-4), tmax, a, b, c)
            if r is not None:
                best = (ti, *r)
                tmax = r[0] - EPS
        depth = layers[1, py, px, 3]
        if best is None:
            assert depth == 0
            continue
        ti, t, u, v = best
        assert F(t).view(np.uint32) == depth.view(np.uint32)
        l0, l1, l2 = (F(1.0) - u) - v, u, v
        na, nb, nc = (verts[tris[ti, k], 4:7] for k in range(3))
        nn = (na * l0 + nb * l1) + nc * l2
        nn = nn * (F(1.0) / np.sqrt(_dot(nn, nn)))
        assert np.array_equal(nn.view(np.uint32), layers[1, py, px, :3].view(np.uint32))
        checked += 1
    assert checked >= 3
