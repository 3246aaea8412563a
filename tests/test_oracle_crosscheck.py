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


# ---------------------------------------------------------------------------------------------------------------
# A whole path and a whole reconstruction block, restated a second time in numpy float32 from the GLSL
# (render.glsl:81-147 with next-event estimation, the diffuse / mirror / dielectric branches of material.glsl and
# Russian roulette; reconstruction.glsl:22-66), for triangle scenes.  Shared with the oracle are only the things the
# GLSL leaves to the driver and orc_math.h therefore FIXES (sin, cos, exp: evaluated through orc_math_eval, and
# normalize(v) = v * (1 / sqrt(dot(v, v)))); every control-flow decision, RNG draw and fp32 operation order below is
# written from the shader text, not from the C++.
class _Rng:
    def __init__(self, seed):
        s = np.uint32(seed)
        with np.errstate(over="ignore"):
            s = (s ^ np.uint32(61)) ^ (s >> np.uint32(16))
            s = s * np.uint32(9)
            s = s ^ (s >> np.uint32(4))
            s = s * np.uint32(0x27d4eb2d)
            s = s ^ (s >> np.uint32(15))
        self.s = s

    def uniform(self):  # rand.glsl:2-7,18-20
        x = self.s
        x = x ^ (x << np.uint32(13))
        x = x ^ (x >> np.uint32(17))
        x = x ^ (x << np.uint32(5))
        self.s = x
        return F(x) * F(1.0 / 4294967296.0)


def _m(oracle, fn, x):
    return _libs.math_eval(oracle.orc_math_eval, fn, np.array([x], F))[0]


def _normalize(v):
    return v * (F(1.0) / np.sqrt(_dot(v, v)))


def _length(v):
    return np.sqrt(_dot(v, v))


def _reflect(i, n):  # GLSL reflect: I - 2 * dot(N, I) * N
    return i - (F(2.0) * _dot(n, i)) * n


class _NumpyTracer:
    def __init__(self, oracle, scene, eps=EPS):
        self.o, self.eps = oracle, eps
        self.tris, self.verts = scene.array("triangles"), scene.array("vertices")
        self.mats, self.emitters = scene.array("materials").reshape(-1), scene.array("emitters")
        self.diffuse, self.emissive, self.dielectric = scene.array("diffuse"), scene.array("emissive"), scene.array("dielectric")
        assert scene.info.num_spheres == 0 and scene.info.num_quads == 0
        self.n_ext = self.n_shadow = 0

    def intersect(self, o, d, tmin, tmax):  # scene.glsl:134-157 (triangles only) + :160-175
        best = None
        for ti in range(self.tris.shape[0]):
            a, b, c = (self.verts[self.tris[ti, k], :3] for k in range(3))
            r = _triangle(o, d, tmin, tmax, a, b, c)
            if r is not None:
                best = (ti, *r)
                tmax = r[0] - self.eps
        return best

    def populate(self, o, d, hit):  # scene.glsl:164 + shapes/triangle.glsl:54-78
        ti, t, u, v = hit
        p = o + t * d
        l0, l1, l2 = (F(1.0) - u) - v, u, v
        na, nb, nc = (self.verts[self.tris[ti, k], 4:7] for k in range(3))
        n = _normalize((na * l0 + nb * l1) + nc * l2)
        bt = np.array([0, 1, 0], F) if abs(n[0]) > abs(n[1]) else np.array([1, 0, 0], F)
        tg = _normalize(_cross(n, bt))
        bt = _cross(n, tg)
        return p, n, (tg, bt, n)

    def sample_emitter(self, rng, ref):  # scene.glsl:54-89 + shapes/triangle.glsl:81-102 + rand.glsl:42-50
        sample = rng.uniform()
        emitter = 0
        for i in range(self.emitters.shape[0]):
            sample = sample - self.emitters[i, 1]
            if sample < 0:
                emitter = i
                break
        shape = int(self.emitters.view(np.uint32)[emitter, 0])
        a, b, c = (self.verts[self.tris[shape, k]] for k in range(3))
        nn = _cross(b[:3] - a[:3], c[:3] - a[:3])
        area = _length(nn) / F(2.0)
        u, v = rng.uniform(), rng.uniform()
        if u + v > F(1.0):
            u = F(1.0) - v
            v = F(1.0) - u  # (sic) rand.glsl:46-47, SURVEY Q5
        lam = (u, v, (F(1.0) - u) - v)
        sn = _normalize((a[4:7] * lam[0] + b[4:7] * lam[1]) + c[4:7] * lam[2])
        sp = (a[:3] * lam[0] + b[:3] * lam[1]) + c[:3] * lam[2]
        spdf = F(1.0) / area
        power = self.emissive[self.mats[shape] & 0xFFFFFF, :3]
        direction = sp - ref
        dist = _length(direction)
        direction = direction / dist
        shadow = (ref, direction, F(2.0) * self.eps, dist - self.eps)
        cos_theta = -_dot(direction, sn)
        if cos_theta < 0:
            return np.zeros(3, F), shadow
        pdf = ((self.emitters[emitter, 1] * spdf) * dist) * dist / cos_theta
        return power / pdf, shadow

    def path(self, rng, o, d, max_bounces, rr_start=3):  # render.glsl:81-147
        pi = F(3.1415926535897932384626433832795)
        total, throughput, ext = np.zeros(3, F), np.ones(3, F), np.zeros(3, F)
        was_discrete, tmin, tmax = True, self.eps, F(np.inf)
        log = []
        for bounce in range(max_bounces):
            self.n_ext += 1
            hit = self.intersect(o, d, tmin, tmax)
            if hit is None:
                log.append((-1, F(0), int(rng.s), throughput.copy(), total.copy(), 0))
                return log
            p, n, frame = self.populate(o, d, hit)
            mat = int(self.mats[hit[0]])
            tag, idx = mat >> 24, mat & 0xFFFFFF
            dist = _length(o - p)
            e = -ext * dist
            throughput = throughput * np.array([_m(self.o, 3, x) for x in e], F)
            if tag == _abi.MAT_EMISSIVE and was_discrete:
                total = total + throughput * self.emissive[idx, :3]
            shadow_state = 0
            if tag == _abi.MAT_DIFFUSE:
                importance, (so, sd, smin, smax) = self.sample_emitter(rng, p)
                if _length(importance) > self.eps and _dot(sd, n) > 0:
                    self.n_shadow += 1
                    if self.intersect(so, sd, smin, smax) is None:
                        bsdf = (_dot(n, sd) * self.diffuse[idx, :3]) / pi  # material.glsl:21-23
                        total = total + (throughput * bsdf) * importance
                        shadow_state = 2
                    else:
                        shadow_state = 1
            wo_written = True
            if tag == _abi.MAT_DIFFUSE:  # rand.glsl:22-30, material.glsl:38-43
                u, v = rng.uniform(), rng.uniform()
                r = np.sqrt(u)
                theta = (F(2.0) * pi) * v
                local = np.array([r * _m(self.o, 1, theta), r * _m(self.o, 0, theta), np.sqrt(max(F(0), F(1.0) - u))], F)
                wo = (frame[0] * local[0] + frame[1] * local[1]) + frame[2] * local[2]
                weight = self.diffuse[idx, :3]
            elif tag == _abi.MAT_MIRROR:
                wo, weight = _reflect(d, n), np.ones(3, F)
            elif tag == _abi.MAT_DIELECTRIC:  # material.glsl:52-87
                eta = self.dielectric[idx, 3]
                eta_inv = F(1.0) / eta
                cos_i = -_dot(n, d)
                normal = n
                inside = cos_i > 0
                if cos_i < 0:
                    eta = eta_inv
                    eta_inv = F(1.0) / eta
                    normal = -normal
                    cos_i = -cos_i
                k = F(1.0) - (eta_inv * eta_inv) * (F(1.0) - cos_i * cos_i)
                if k <= 0:
                    wo = _reflect(d, normal)
                else:
                    cos_o = np.sqrt(k)
                    rho_par = (eta * cos_i - cos_o) / (eta * cos_i + cos_o)
                    rho_orth = (cos_i - eta * cos_o) / (cos_i + eta * cos_o)
                    f_r = F(0.5) * (rho_par * rho_par + rho_orth * rho_orth)
                    if rng.uniform() < f_r:
                        wo = _reflect(d, normal)
                    else:
                        inside = not inside
                        parallel = d - _dot(d, normal) * normal
                        wo = eta_inv * parallel - np.sqrt(k) * normal
                if inside:
                    ext = self.dielectric[idx, :3].copy()
                weight = np.ones(3, F)
            else:  # emissive: wo is never written, the weight is 0
                wo, weight, wo_written = np.zeros(3, F), np.zeros(3, F), False
            throughput = throughput * weight
            o, d, tmin, tmax = p, wo, F(2.0) * self.eps, F(np.inf)
            was_discrete = tag != _abi.MAT_DIFFUSE
            terminate = False
            if bounce > rr_start:
                q = min(F(0.99), max(throughput[0], max(throughput[1], throughput[2])))
                if rng.uniform() > q:
                    terminate = True
                else:
                    throughput = throughput / q
            log.append((hit[0], hit[1], int(rng.s), throughput.copy(), total.copy(), shadow_state))
            if terminate or not wo_written:
                break
        return log


def _small_material_scene():
    """A closed box of 12 diffuse triangles with an emissive ceiling panel, a mirror triangle and a tinted glass
    slab: every branch of sampleBSDF a triangle scene can reach, cheap enough for a Python linear scan."""
    D, MI, DI, EM = _abi.MAT_DIFFUSE, _abi.MAT_MIRROR, _abi.MAT_DIELECTRIC, _abi.MAT_EMISSIVE
    verts, tris, mats = [], [], []

    def quad(p0, e1, e2, n, mat):
        base = len(verts)
        p0, e1, e2 = np.array(p0, float), np.array(e1, float), np.array(e2, float)
        for k, q in enumerate((p0, p0 + e1, p0 + e1 + e2, p0 + e2)):
            verts.append((*q, (k in (1, 2)) * 1.0, *n, (k in (2, 3)) * 1.0))
        tris.extend([(base, base + 1, base + 2), (base, base + 2, base + 3)])
        mats.extend([mat, mat])

    quad((-1, 0, -1), (2, 0, 0), (0, 0, 2), (0, 1, 0), (D, 0))      # floor
    quad((-1, 2, -1), (0, 0, 2), (2, 0, 0), (0, -1, 0), (D, 0))     # ceiling
    quad((-1, 0, -1), (0, 2, 0), (2, 0, 0), (0, 0, 1), (D, 1))      # back
    quad((-1, 0, -1), (0, 0, 2), (0, 2, 0), (1, 0, 0), (D, 2))      # left
    quad((1, 0, -1), (0, 2, 0), (0, 0, 2), (-1, 0, 0), (D, 1))      # right
    quad((-0.4, 1.98, -0.4), (0, 0, 0.8), (0.8, 0, 0), (0, -1, 0), (EM, 0))  # light
    quad((-0.9, 0.1, -0.6), (0.7, 0, 0.5), (0, 0.9, 0), (-0.58, 0, 0.81), (MI, 0))  # mirror, facing the room
    quad((0.1, 0.2, 0.2), (0.7, 0, 0), (0, 0.8, 0), (0, 0, 1), (DI, 0))   # glass pane (front face)
    quad((0.1, 0.2, 0.05), (0, 0.8, 0), (0.7, 0, 0), (0, 0, -1), (DI, 0))  # glass pane (back face)
    half = np.deg2rad(-8.0) / 2
    cam = ((0.0, 1.0, 3.4), (float(np.sin(half)), 0.0, 0.0, float(np.cos(half))), 42.0)
    return _libs.CustomScene(cam, triangles=tris, vertices=verts, materials=mats,
                             diffuse=[(0.7, 0.7, 0.7, 0), (0.3, 0.6, 0.3, 0), (0.7, 0.25, 0.2, 0)],
                             dielectric=[(0.8, 0.2, 0.4, 1.5)], emissive=[(14, 13, 12, 0)])


def test_whole_paths_agree_with_numpy_restatement(oracle, hosttest):
    """render.glsl:81-147 bounce by bounce: hit id, t, RNG state after the bounce, throughput, accumulated radiance and
    the shadow-ray outcome of every vertex of every path of a 14x10 frame, bit for bit, plus the ray counts."""
    scene = _small_material_scene()
    w, h, max_bounces = 14, 10, 12
    blocks = _libs.generate_blocks(hosttest, w, h, 1, block_size=64)
    blk = blocks[0]
    op = _libs.orc_params(max_bounces=max_bounces, use_bvh=0, block_size=64)
    tracer = _NumpyTracer(oracle, scene)
    one = np.zeros(1, dtype=_abi.RAY_DTYPE)
    seen_tags, deepest, shadow_states = set(), 0, set()
    for py in range(h):
        for px in range(w):
            logbuf = (_libs.OrcPathVertex * 64)()
            n = oracle.orc_trace_path(C.byref(scene.view), _libs.ptr(blocks), px, py, C.byref(op), logbuf, 64)
            oracle.orc_camera_ray(scene.view.scene.ptr, float(F(px) + blk["sample_offset"][0]),
                                  float(F(py) + blk["sample_offset"][1]), float(w), float(h), 1e-4, _libs.ptr(one))
            rng = _Rng(int(blk["seed"]) + px + py * int(blk["dimension"][0]))  # render.glsl:156
            mine = tracer.path(rng, one[0]["origin"].copy(), one[0]["direction"].copy(), max_bounces)
            assert n == len(mine), (px, py, n, len(mine))
            deepest = max(deepest, n)
            for k, (sid, t, state, thr, tot, sh) in enumerate(mine):
                pv = logbuf[k]
                where = (px, py, k)
                assert pv.shape_id == sid, where
                assert F(pv.t).view(np.uint32) == F(t).view(np.uint32), where
                assert pv.rng_after == state, where
                assert np.array_equal(np.array(pv.throughput[:], F).view(np.uint32), thr.view(np.uint32)), where
                assert np.array_equal(np.array(pv.total[:], F).view(np.uint32), tot.view(np.uint32)), where
                assert pv.shadow_state == sh, where
                shadow_states.add(sh)
                if sid >= 0:
                    seen_tags.add(int(scene.array("materials")[sid]) >> 24)
    assert seen_tags == {_abi.MAT_DIFFUSE, _abi.MAT_MIRROR, _abi.MAT_DIELECTRIC, _abi.MAT_EMISSIVE}
    assert shadow_states == {0, 1, 2} and deepest >= 6  # roulette region reached (bounce > 3)
    # the frame's ray counts: the numpy tracer counted its own intersectScene calls
    acc = np.zeros((h, w, 4), F)
    st = _libs.OrcStats()
    assert oracle.orc_render(C.byref(scene.view), _libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(acc), C.byref(st),
                             0) == 0
    assert (st.n_extension_rays, st.n_shadow_rays) == (tracer.n_ext, tracer.n_shadow)


def test_reconstruction_block_agrees_with_numpy_restatement(oracle):
    """reconstruction.glsl:22-66 on a 16x16 frame made of four 8x8 blocks (so aprons, block edges and image edges all
    occur), R = 2, sigma = 0.5, with a NaN sample: written from the shader, compared bit for bit with
    orc_reconstruct_frame.  exp() goes through the oracle's two specifications (orc_math.h: fn 3 for the spatial
    weight, fn 6 for the bilateral one), as the GLSL leaves it to the driver."""
    rng = np.random.default_rng(31)
    w = h = 16
    bs, R, sigma = 8, 2, F(0.5)
    blocks = np.zeros(4, dtype=_abi.BLOCK_DTYPE)
    so = rng.random(2).astype(F)
    for k, (bx, by) in enumerate([(0, 0), (8, 0), (0, 8), (8, 8)]):
        blocks[k] = (k, 100 + k, (bx, by), (bs, bs), (w, h), so)
    rad = np.exp(rng.standard_normal((h, w, 4))).astype(F)
    rad[..., 3] = 1.0
    rad[5, 9, 2] = np.nan
    nrm = rng.standard_normal((h, w, 4)).astype(F)
    nrm[..., :3] /= np.linalg.norm(nrm[..., :3], axis=2, keepdims=True)
    nrm[:6, :6, :3] = (0, 0, 1)
    acc_o = (rng.random((h, w, 4)) * 0.1).astype(F)  # the pass ADDS to what is there
    want = acc_o.copy()
    op = _libs.orc_params(block_size=bs, radius=R, stddev=float(sigma))
    assert oracle.orc_reconstruct_frame(_libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(rad), _libs.ptr(nrm), None,
                                        _libs.ptr(acc_o), 0) == 0
    gauss = F(-1.0) / ((F(2.0) * sigma) * sigma)
    curve = _m(oracle, 3, (gauss * F(R)) * F(R))
    for blk in blocks:
        ox, oy = (int(v) for v in blk["origin"])
        dim = int(blk["dimension"][0])

        def load(layer, lx, ly):  # the block's own intermediate texture: zeros outside it
            if 0 <= lx < dim and 0 <= ly < dim:
                return (rad if layer == 0 else nrm)[oy + ly, ox + lx]
            return np.zeros(4, F)

        for gy_ in range(-R, dim + R):      # gl_GlobalInvocationID - R: the dispatch covers the block and its apron
            for gx_ in range(-R, dim + R):
                X, Y = ox + gx_, oy + gy_
                if not (0 <= X < w and 0 <= Y < h):
                    continue  # robust image access: stores outside the image are dropped
                out = want[Y, X].copy()
                nc = load(1, gx_, gy_)[:3]
                for dx in range(-R, R + 1):
                    if gx_ + dx < 0 or gx_ + dx >= dim:
                        continue
                    for dy in range(-R, R + 1):
                        if gy_ + dy < 0 or gy_ + dy >= dim:
                            continue
                        sx, sy = (F(dx) + so[0]) - F(0.5), (F(dy) + so[1]) - F(0.5)
                        wgt = _m(oracle, 3, gauss * (sx * sx + sy * sy)) - curve
                        if wgt < 0:
                            continue
                        cw = load(0, gx_ + dx, gy_ + dy)
                        no = load(1, gx_ + dx, gy_ + dy)[:3] - nc
                        wgt = wgt * _m(oracle, 6, _dot(no, no) * F(2.0) + F(0.0))
                        weighted = wgt * cw
                        if np.isnan(weighted).any():
                            continue
                        out = out + weighted
                want[Y, X] = out
    assert np.array_equal(want.view(np.uint32), acc_o.view(np.uint32))
