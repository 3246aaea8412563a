"""The oracle's mode 3 (the reference's linear scan over the primitives whose box the ray pierces,
oracle/hijiki_oracle.cpp "mode 3") against the plain scans it stands in for at BASELINE sizes: mode 0
(scene.glsl:134-157, the reference default) and mode 2 (the same without the >100-shape failsafe).  Same ids,
t, uv and tie flags on every ray, same frames bit for bit, same ray counts."""
import ctypes as C

import numpy as np
import pytest

import _libs
from hijiki_b200 import _abi


def _scene(hosttest, kind):
    if kind == "cbox":
        return _libs.HostScene.from_obj(hosttest, _libs.CBOX_OBJ, put_spheres=False, with_bvh2=False), 0
    if kind == "cbox_spheres":
        return _libs.HostScene.from_obj(hosttest, _libs.CBOX_OBJ, put_spheres=True, with_bvh2=False), 0
    if kind == "spheres":
        return _libs.HostScene.spheres(hosttest, 5, with_bvh2=False), 2
    if kind == "terrain":
        return _libs.HostScene.terrain(hosttest, 24, with_bvh2=False), 2
    if kind == "quad_room":
        return _libs.quad_room_scene(), 0
    raise KeyError(kind)


def _rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    v = scene.array("vertices")[:, :3]
    lo, hi = v.min(axis=0) - 0.5, v.max(axis=0) + 0.5
    rays = np.zeros(n, dtype=_abi.RAY_DTYPE)
    rays["origin"] = (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["direction"] = d.astype(np.float32)
    rays["t_min"] = 1e-4
    rays["t_max"] = np.inf
    # a third of the rays: finite tMax (shadow-ray like); a few: directions off unit length, origins far outside
    third = n // 3
    rays["t_max"][:third] = (rng.random(third) * 4).astype(np.float32)
    rays["direction"][third:third + 200] *= (10.0 ** rng.uniform(-2, 2, (200, 1))).astype(np.float32)
    rays["origin"][third + 200:third + 300] *= np.float32(30.0)
    rays["t_min"][third + 300:third + 350] = -1.0
    return rays


def _trace(O, scene, rays, mode):
    n = rays.size
    ids, t, uv, tie = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros((n, 2), np.float32), np.zeros(n, np.uint8)
    assert O.orc_trace(C.byref(scene.view), _libs.ptr(rays), n, mode, 1e-4, _libs.ptr(ids), _libs.ptr(t), _libs.ptr(uv),
                       _libs.ptr(tie), 0) == 0
    return ids, t, uv, tie


@pytest.mark.parametrize("kind", ["cbox", "cbox_spheres", "spheres", "terrain", "quad_room"])
def test_culled_scan_gives_the_plain_scan_on_every_ray(oracle, hosttest, kind):
    scene, plain = _scene(hosttest, kind)
    n_cam = (48, 32) if kind.startswith("cbox") else (96, 64)
    rays = np.concatenate([_libs.camera_rays(scene, *n_cam), _rays(scene, 1500 if kind.startswith("cbox") else 6000, 3)])
    a = _trace(oracle, scene, rays, plain)
    b = _trace(oracle, scene, rays, 3)
    finite = np.isfinite(a[1])  # t = +inf artefact hits of exactly perpendicular directions are not reproduced
    assert finite.mean() > 0.99
    assert np.array_equal(a[0][finite], b[0][finite])
    assert np.array_equal(a[1][finite].view(np.uint32), b[1][finite].view(np.uint32))
    assert np.array_equal(a[2][finite].view(np.uint32), b[2][finite].view(np.uint32))
    assert np.array_equal(a[3][finite], b[3][finite])
    assert (a[0] >= 0).mean() > 0.3
    occ_a, occ_b = np.zeros(rays.size, np.uint8), np.zeros(rays.size, np.uint8)
    assert oracle.orc_occluded(C.byref(scene.view), _libs.ptr(rays), rays.size, plain, 1e-4, _libs.ptr(occ_a), 0) == 0
    assert oracle.orc_occluded(C.byref(scene.view), _libs.ptr(rays), rays.size, 3, 1e-4, _libs.ptr(occ_b), 0) == 0
    assert np.array_equal(occ_a[finite], occ_b[finite])


@pytest.mark.parametrize("kind,w,h,spp,bounces", [("cbox", 40, 28, 1, 8), ("cbox_spheres", 36, 24, 1, 1000),
                                                  ("spheres", 96, 64, 2, 16), ("terrain", 96, 64, 2, 8),
                                                  ("quad_room", 64, 48, 2, 12)])
def test_culled_scan_renders_the_same_frame(oracle, hosttest, kind, w, h, spp, bounces):
    scene, plain = _scene(hosttest, kind)
    blocks = _libs.generate_blocks(hosttest, w, h, spp, 64)
    out = []
    for mode in (plain, 3):
        acc = np.zeros((h, w, 4), np.float32)
        st = _libs.OrcStats()
        op = _libs.orc_params(max_bounces=bounces, use_bvh=mode, block_size=64)
        assert oracle.orc_render(C.byref(scene.view), _libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(acc),
                                 C.byref(st), 0) == 0
        out.append((acc, (st.n_paths, st.n_extension_rays, st.n_shadow_rays)))
    assert out[0][1] == out[1][1]
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32))
