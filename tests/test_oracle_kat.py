"""Known-answer tests that pin the oracle to the reference arithmetic.

The reference ships no tests or golden vectors; these values were derived by hand from
reference shader/rand.glsl:1-20, render.glsl:26-36 and reconstruction.glsl:29-46.
"""
import ctypes as C

import numpy as np
import pytest

import _libs
from hijiki_b200 import _abi

RNG_KAT = [
    (0x00000000, 0xC0A9496A, [0xD90BC8A8, 0xA3CD8C47, 0x5AE9C9C5, 0x19FA5D8D]),
    (0x00000001, 0x27922C9D, [0x22360E3D, 0x9DCA2765, 0xFDFB9536, 0x64FF4198]),
    (0x00000002, 0xC6793575, [0xFA2B46DE, 0xCCE93B66, 0x9B345A24, 0x1E6A919C]),
    (0x0000003D, 0x00000000, [0, 0, 0, 0]),  # seed 61 hashes to the absorbing state
    (0x00003039, 0x0DDEEC13, [0xDBC0639D, 0x21C6A0C4, 0x4E151F4B, 0x527D3F15]),
    (0xDEADBEEF, 0x572E7C2D, [0x8DD99F78, 0x78EECC03, 0x8CB16A34, 0x9F00E32F]),
    (0xFFFFFFFF, 0x70F499D3, [0x9A1F8EB4, 0x12EE5150, 0xC0439B72, 0xD69DDE64]),
]


@pytest.mark.parametrize("seed,state,draws", RNG_KAT)
def test_rng_known_answers(oracle, seed, state, draws):
    assert oracle.orc_seed_rng(seed) == state
    s = C.c_uint32(state)
    assert [oracle.orc_rand_uint(C.byref(s)) for _ in range(4)] == draws


def test_rng_floats(oracle):
    s = C.c_uint32(oracle.orc_seed_rng(0))
    f = [oracle.orc_rand_uniform_float(C.byref(s)) for _ in range(2)]
    assert f[0] == pytest.approx(0.847836077, abs=1e-8) and f[1] == pytest.approx(0.639855146, abs=1e-8)
    s = C.c_uint32(oracle.orc_seed_rng(1))
    f = [oracle.orc_rand_uniform_float(C.byref(s)) for _ in range(2)]
    assert f[0] == pytest.approx(0.133637324, abs=1e-8) and f[1] == pytest.approx(0.61636585, abs=1e-8)


def _float_of_next(oracle, wanted_uint):
    # invert one xorshift step by brute force is unnecessary: feed the state whose NEXT output is wanted
    # via the algebraic inverse of xorshift32 (13,17,5)
    def unshift_left(v, s):
        r = v
        for _ in range(32 // s + 1):
            r = v ^ ((r << s) & 0xFFFFFFFF)
        return r

    def unshift_right(v, s):
        r = v
        for _ in range(32 // s + 1):
            r = v ^ (r >> s)
        return r

    st = unshift_left(wanted_uint, 5)
    st = unshift_right(st, 17)
    st = unshift_left(st, 13)
    s = C.c_uint32(st)
    f = oracle.orc_rand_uniform_float(C.byref(s))
    assert s.value == wanted_uint
    return f


def test_uniform_float_can_return_one(oracle):
    """float(uint) * 2^-32 rounds to nearest: 0xFFFFFF80 and above give exactly 1.0 (rand.glsl:18-20)."""
    assert _float_of_next(oracle, 0xFFFFFF7F) == np.float32(0.99999994)
    assert _float_of_next(oracle, 0xFFFFFF80) == np.float32(1.0)
    assert _float_of_next(oracle, 0xFFFFFFFF) == np.float32(1.0)


CAMERA_KAT = [
    (800, 600, 0.0, 0.0, (-0.23561369, 0.15247145, -0.95981175)),
    (800, 600, 400.0, 300.0, (0.0, -0.02530457, -0.9996798)),
    (800, 600, 799.5, 599.5, (0.23534773, -0.20056581, -0.95099145)),
    (1920, 1080, 960.5, 540.5, (1.2841e-4, -0.02543294, -0.99967659)),
]


@pytest.mark.parametrize("w,h,px,py,expect", CAMERA_KAT)
def test_camera_known_answers(oracle, cbox, w, h, px, py, expect):
    ray = np.zeros(1, dtype=_abi.RAY_DTYPE)
    oracle.orc_camera_ray(cbox.view.scene.ptr, px, py, float(w), float(h), 1e-4, _libs.ptr(ray))
    assert np.allclose(ray[0]["direction"], expect, atol=2e-6)
    assert np.allclose(ray[0]["origin"], (0.0, 0.91, 5.41), atol=1e-6)
    assert ray[0]["t_min"] == np.float32(1e-4) and np.isinf(ray[0]["t_max"])
    assert abs(np.linalg.norm(ray[0]["direction"].astype(np.float64)) - 1.0) < 1e-6


def test_recon_spatial_weights(oracle):
    """reconstruction.glsl:29-30,43-46 with R=2, sigma=0.5."""
    w = np.zeros(25, dtype=np.float32)
    oracle.orc_recon_spatial_weights(2, 0.5, 0.5, 0.5, _libs.ptr(w))
    assert (w >= 0).sum() == 13  # the four taps at distance exactly R have weight exp(-8) - exp(-8) = 0
    pos = w[w > 0]
    assert pos.size == 9
    assert pos.max() == pytest.approx(0.99966, abs=2e-5)
    assert np.sort(pos)[0] == pytest.approx(0.01798, abs=2e-5)
    assert np.sort(pos)[4] == pytest.approx(0.13500, abs=2e-5)
    assert pos.sum() == pytest.approx(1.61158, abs=1e-4)
    oracle.orc_recon_spatial_weights(2, 0.5, 0.0, 0.0, _libs.ptr(w))
    pos = w[w > 0]
    assert pos.size == 12
    assert np.sort(pos)[-1] == pytest.approx(0.36754, abs=2e-5)
    assert np.sort(pos)[0] == pytest.approx(0.0064, abs=1e-4)
    assert pos.sum() == pytest.approx(1.52140, abs=1e-4)


def test_cbox_fixture_facts(cbox):
    """SURVEY.md §8-L 'cbox facts', derived from parsing the OBJ the way Scene::from_obj does."""
    info = cbox.info
    assert (info.num_spheres, info.num_quads, info.num_triangles, info.num_emitters) == (0, 0, 6332, 2)
    assert cbox.array("vertices").shape[0] == 3668
    assert cbox.array("bvh").shape[0] == 12663  # 2n-1 nodes, one shape per leaf
    mats = cbox.array("materials")[:, 0]
    tags = mats >> 24
    assert list(np.nonzero(tags == _abi.MAT_EMISSIVE)[0]) == [6324, 6325]
    assert (tags[:6320] == _abi.MAT_DIFFUSE).all()
    em = cbox.array("emitters")
    assert em.view(np.uint32)[:, 0].tolist() == [6324, 6325] and np.allclose(em[:, 1], 0.5)
    assert np.allclose(cbox.array("emissive")[0, :3], 15.0)
    v = cbox.array("vertices")
    assert np.allclose(v[:, :3].min(axis=0), (-1.0, 0.0, -1.04), atol=1e-5)
    assert np.allclose(v[:, :3].max(axis=0), (1.0, 1.59, 0.99), atol=1e-5)
    cam = info.camera
    assert np.allclose(list(cam.position)[:3], (0, 0.91, 5.41))
    assert np.allclose(list(cam.rotation), (-0.0126533, 0, 0, 0.99991995), atol=1e-6)
    assert cam.fov == pytest.approx(27.7)
    # root skip pointer sentinel (src/main.rs:231)
    assert cbox.array("bvh").view(np.uint32)[0, 7] == 1000000


def test_put_cbox_spheres(cbox_spheres):
    """main()'s --put-cbox-spheres block (src/main.rs:1463-1483): a mirror and a checkerboard sphere."""
    info = cbox_spheres.info
    assert info.num_spheres == 2 and info.num_triangles == 6332
    tags = cbox_spheres.array("materials")[:2, 0] >> 24
    assert sorted(tags.tolist()) == [_abi.MAT_DIFFUSECBOARD, _abi.MAT_MIRROR]
    assert np.allclose(cbox_spheres.array("spheres")[:, 3], 0.3263)


def test_linear_and_bvh2_modes_agree_off_ties(oracle, cbox):
    """scene.glsl USE_BVH=0 vs USE_BVH=1 report the same primitive except on ties."""
    rays = _libs.camera_rays(cbox, 96, 72)
    n = rays.size
    out = {}
    for mode in (0, 1):
        ids = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        tie = np.zeros(n, np.uint8)
        oracle.orc_trace(C.byref(cbox.view), _libs.ptr(rays), n, mode, 1e-4, _libs.ptr(ids), _libs.ptr(t), None,
                         _libs.ptr(tie) if mode == 0 else None, 0)
        out[mode] = (ids, t, tie)
    keep = out[0][2] == 0
    assert (out[0][0][keep] == out[1][0][keep]).all()
    assert (out[0][1][keep].view(np.uint32) == out[1][1][keep].view(np.uint32)).all()
    assert (out[0][0] >= 0).sum() > 0.8 * n
