"""Kernel LOGIC vs the oracle, in the CPU-only container.

tests/native compiles the product's device headers (traverse.cuh, shade.cuh, recon.cuh — the code
the CUDA kernels run) for the host.  Same seeded inputs through that code and through the oracle
(the literal restatement of the reference GLSL) must agree bit for bit, ties excepted.
The GPU tests (test_gpu_parity.py) repeat these through the C ABI on the device.
"""
import ctypes as C

import numpy as np
import pytest

import _libs
from hijiki_b200 import _abi


def _harness(hosttest, scene, pad=1e-5):
    err = C.create_string_buffer(256)
    h = hosttest.ht_create(C.byref(scene.view), pad, err, 256)
    assert h, err.value
    return h


def _random_rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    v = scene.array("vertices")[:, :3]
    lo, hi = v.min(axis=0) - 0.5, v.max(axis=0) + 0.5
    rays = np.zeros(n, dtype=_abi.RAY_DTYPE)
    rays["origin"] = (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["direction"] = d.astype(np.float32)
    rays["t_min"] = 2e-4
    rays["t_max"] = np.inf
    # a few axis-parallel and degenerate directions (zero components -> 1/0 in scene.glsl:100)
    rays["direction"][:6] = np.eye(3, dtype=np.float32).repeat(2, axis=0) * np.array([1, -1] * 3)[:, None]
    return rays


def _trace_oracle(oracle, scene, rays, mode=0, want_tie=True):
    n = rays.size
    ids, t, uv = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros((n, 2), np.float32)
    tie = np.zeros(n, np.uint8)
    rc = oracle.orc_trace(C.byref(scene.view), _libs.ptr(rays), n, mode, 1e-4, _libs.ptr(ids), _libs.ptr(t),
                          _libs.ptr(uv), _libs.ptr(tie) if want_tie else None, 0)
    assert rc == 0
    return ids, t, uv, tie


def _trace_harness(hosttest, h, rays, any_hit=False, chaos=False):
    n = rays.size
    ids, t, uv = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros((n, 2), np.float32)
    hosttest.ht_trace(h, _libs.ptr(rays), n, int(any_hit) | (2 if chaos else 0), 1e-4, _libs.ptr(ids), _libs.ptr(t),
                      _libs.ptr(uv))
    return ids, t, uv


def test_traversal_is_schedule_independent(hosttest, cbox_spheres):
    """Yielding (dynamic fetch) and primitive postponing only reorder work, and the closest-hit rule
    (smallest t, equal t by the lower shape id) is a function of the hit set: same ids, t, uv on every
    ray whatever the schedule — ties included; occlusion is identical everywhere."""
    h = _harness(hosttest, cbox_spheres)
    rays = np.concatenate([_libs.camera_rays(cbox_spheres, 96, 72), _random_rays(cbox_spheres, 20000, 17),
                           _tie_heavy_rays()])
    ids_a, t_a, uv_a = _trace_harness(hosttest, h, rays)
    ids_b, t_b, uv_b = _trace_harness(hosttest, h, rays, chaos=True)
    assert np.array_equal(ids_a, ids_b)
    assert np.array_equal(t_a.view(np.uint32), t_b.view(np.uint32))
    assert np.array_equal(uv_a.view(np.uint32), uv_b.view(np.uint32))
    occ_a, _, _ = _trace_harness(hosttest, h, rays, any_hit=True)
    occ_b, _, _ = _trace_harness(hosttest, h, rays, any_hit=True, chaos=True)
    assert np.array_equal(occ_a, occ_b)
    hosttest.ht_destroy(h)


@pytest.mark.parametrize("which", ["cbox", "cbox_spheres"])
def test_first_hit_ids_bit_exact(oracle, hosttest, request, which):
    """North-star correctness bar: first-hit primitive ids bit-exact, ties excluded and counted."""
    scene = request.getfixturevalue(which)
    h = _harness(hosttest, scene)
    rays = np.concatenate([_libs.camera_rays(scene, 160, 120), _random_rays(scene, 20000, 5)])
    ids_o, t_o, uv_o, tie = _trace_oracle(oracle, scene, rays)
    ids_h, t_h, uv_h = _trace_harness(hosttest, h, rays)
    keep = tie == 0
    assert tie.sum() < 0.002 * rays.size
    assert (ids_o[keep] == ids_h[keep]).all()
    hit = keep & (ids_o >= 0)
    assert hit.sum() > 0.5 * rays.size
    assert (t_o[hit].view(np.uint32) == t_h[hit].view(np.uint32)).all()
    tri = hit & (ids_o >= scene.info.num_spheres)
    assert (uv_o[tri].view(np.uint32) == uv_h[tri].view(np.uint32)).all()
    hosttest.ht_destroy(h)


def test_shadow_rays_exact(oracle, hosttest, cbox):
    """Any-hit == the reference's 'closest hit then test for a hit' shadow overload (scene.glsl:92-96)."""
    h = _harness(hosttest, cbox)
    rays = _random_rays(cbox, 30000, 9)
    rng = np.random.default_rng(1)
    rays["t_max"] = (rng.random(rays.size) * 3).astype(np.float32)
    occ = np.zeros(rays.size, np.uint8)
    assert oracle.orc_occluded(C.byref(cbox.view), _libs.ptr(rays), rays.size, 0, 1e-4, _libs.ptr(occ), 0) == 0
    ids_h, _, _ = _trace_harness(hosttest, h, rays, any_hit=True)
    assert (ids_h == occ).all()
    assert 0.05 < occ.mean() < 0.95
    hosttest.ht_destroy(h)


def test_synthetic_scenes_first_hit(oracle, hosttest):
    for scene in (_libs.HostScene.spheres(hosttest, 4), _libs.HostScene.terrain(hosttest, 48)):
        h = _harness(hosttest, scene)
        rays = _libs.camera_rays(scene, 96, 54)
        ids_o, t_o, _, tie = _trace_oracle(oracle, scene, rays, mode=2)  # linear scan without the >100 failsafe
        ids_h, t_h, _ = _trace_harness(hosttest, h, rays)
        keep = tie == 0
        assert (ids_o[keep] == ids_h[keep]).all()
        hit = keep & (ids_o >= 0)
        assert hit.sum() > 0.3 * rays.size
        assert (t_o[hit].view(np.uint32) == t_h[hit].view(np.uint32)).all()
        hosttest.ht_destroy(h)


def _render_both(oracle, hosttest, scene, w, hgt, spp, bs, max_bounces, use_bvh=0):
    h = _harness(hosttest, scene)
    blocks = _libs.generate_blocks(hosttest, w, hgt, spp, block_size=bs)
    acc_o = np.zeros((hgt, w, 4), np.float32)
    st = _libs.OrcStats()
    op = _libs.orc_params(max_bounces=max_bounces, use_bvh=use_bvh, block_size=bs)
    assert oracle.orc_render(C.byref(scene.view), _libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(acc_o),
                             C.byref(st), 0) == 0
    acc_h = np.zeros((hgt, w, 4), np.float32)
    cnt = np.zeros(3, np.uint64)
    hp = _libs.hjk_params(max_bounces=max_bounces)
    assert hosttest.ht_render(h, _libs.ptr(blocks), blocks.size, C.byref(hp), _libs.ptr(acc_h), None,
                              _libs.ptr(cnt)) == 0
    hosttest.ht_destroy(h)
    return acc_o, st, acc_h, cnt


@pytest.mark.parametrize("which,max_bounces", [("cbox", 8), ("cbox", 1000), ("cbox_spheres", 1000)])
def test_render_accumulator_matches_oracle(oracle, hosttest, request, which, max_bounces):
    """Whole pipeline (raygen, extend, shade, shadow, reconstruction) vs Renderer::render of the oracle:
    same RNG streams, same ray counts, accumulator bit-identical except around tie-affected samples."""
    scene = request.getfixturevalue(which)
    acc_o, st, acc_h, cnt = _render_both(oracle, hosttest, scene, 136, 100, 2, 64, max_bounces)
    assert cnt[0] == st.n_paths == 136 * 100 * 2
    diff = (acc_o.view(np.uint32) != acc_h.view(np.uint32)).any(axis=2)
    assert diff.sum() <= 25 * 2  # at most two tie-affected samples (each touches <= 25 texels)
    assert abs(int(cnt[1]) - int(st.n_extension_rays)) <= 20 and abs(int(cnt[2]) - int(st.n_shadow_rays)) <= 20
    assert np.isfinite(acc_h).all() and acc_h[..., 3].min() > 0


def test_dielectric_mirror_scene_matches_oracle(oracle, hosttest):
    """Config-4 style scene (dielectric + mirror spheres): discrete bounces, Fresnel draw order."""
    scene = _libs.HostScene.spheres(hosttest, 3)
    acc_o, st, acc_h, cnt = _render_both(oracle, hosttest, scene, 96, 64, 2, 64, 16, use_bvh=2)
    diff = (acc_o.view(np.uint32) != acc_h.view(np.uint32)).any(axis=2)
    assert diff.sum() <= 50
    assert abs(int(cnt[1]) - int(st.n_extension_rays)) <= 40


def test_reconstruction_matches_oracle_on_synthetic_layers(oracle, hosttest):
    """reconstruction.glsl on random feature buffers, incl. ragged edge blocks, NaN samples and the
    'last block carries the next pass' offset' quirk."""
    w, hgt, bs = 150, 70, 64
    rng = np.random.default_rng(11)
    blocks = _libs.generate_blocks(hosttest, w, hgt, 2, block_size=bs)[:6]  # one pass; block 5 has another offset
    assert not (blocks[5]["sample_offset"] == blocks[0]["sample_offset"]).all()
    rad = np.exp(rng.standard_normal((hgt, w, 4))).astype(np.float32)
    rad[..., 3] = 1.0
    rad[10, 20, 0] = np.nan
    nrm = rng.standard_normal((hgt, w, 4)).astype(np.float32)
    nrm[..., :3] /= np.linalg.norm(nrm[..., :3], axis=2, keepdims=True)
    alb = rng.random((hgt, w, 4)).astype(np.float32)
    for albedo in (None, alb):
        acc_o = rng.random((hgt, w, 4)).astype(np.float32)
        acc_h = acc_o.copy()
        op = _libs.orc_params(block_size=bs)
        assert oracle.orc_reconstruct_frame(_libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(rad),
                                            _libs.ptr(nrm), _libs.ptr(albedo) if albedo is not None else None,
                                            _libs.ptr(acc_o), 0) == 0
        hp = _libs.hjk_params()
        assert hosttest.ht_denoise(_libs.ptr(blocks), blocks.size, C.byref(hp), _libs.ptr(rad), _libs.ptr(nrm),
                                   _libs.ptr(albedo) if albedo is not None else None, _libs.ptr(acc_h)) == 0
        assert np.array_equal(acc_o.view(np.uint32), acc_h.view(np.uint32))


def test_sum_of_weights_interior_pixel(oracle, hosttest):
    """SURVEY §8c invariant: with equal normals, an interior texel gains the tap-weight sum per pass."""
    w = hgt = 64
    blocks = _libs.generate_blocks(hosttest, w, hgt, 1, block_size=64)
    blocks["sample_offset"] = (0.5, 0.5)
    rad = np.ones((hgt, w, 4), np.float32)
    nrm = np.zeros((hgt, w, 4), np.float32)
    nrm[..., 2] = 1
    acc = np.zeros((hgt, w, 4), np.float32)
    hp = _libs.hjk_params()
    assert hosttest.ht_denoise(_libs.ptr(blocks), 1, C.byref(hp), _libs.ptr(rad), _libs.ptr(nrm), None,
                               _libs.ptr(acc)) == 0
    assert acc[32, 32, 3] == pytest.approx(1.61158, abs=1e-4)
    assert acc[0, 0, 3] < acc[32, 32, 3]


def _tie_heavy_rays(n=30000, seed=2):
    """Rays dropped onto the floor under and around the teapot, whose base is coplanar with the floor:
    many of them see two surfaces closer than M_EPS to each other (SURVEY Q1 ties)."""
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=_abi.RAY_DTYPE)
    rays["origin"] = np.stack([rng.random(n) * 1.2 - 0.6, rng.random(n) * 0.5 + 0.05, rng.random(n) * 1.2 - 0.6],
                              1).astype(np.float32)
    d = np.stack([rng.standard_normal(n) * 0.05, -np.ones(n), rng.standard_normal(n) * 0.05], 1)
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["direction"][: n // 8] = (0, -1, 0)
    rays["t_min"], rays["t_max"] = 1e-4, np.inf
    return rays


def test_exact_tie_mode_reproduces_the_linear_scan_winner(oracle, hosttest, cbox):
    """HJK_RENDER_EXACT_TIES: among hits closer than M_EPS the reference's winner depends on primitive
    order (scene.glsl:134-157); the exact mode replays that order, so ids/t agree on EVERY ray."""
    h = _harness(hosttest, cbox)
    rays = _tie_heavy_rays()
    ids_o, t_o, uv_o, tie = _trace_oracle(oracle, cbox, rays)
    assert tie.sum() >= 10
    ids_d, t_d, _ = _trace_harness(hosttest, h, rays)
    assert ((ids_d != ids_o) & (tie == 0)).sum() == 0  # default mode: can differ from the linear scan on ties only
    # ... and its rule (closest hit, equal t by the lower id) does not depend on the visiting order
    ids_dc, t_dc, _ = _trace_harness(hosttest, h, rays, chaos=True)
    assert np.array_equal(ids_d, ids_dc) and np.array_equal(t_d.view(np.uint32), t_dc.view(np.uint32))
    hosttest.ht_set_exact(1)
    try:
        ids_e, t_e, uv_e = _trace_harness(hosttest, h, rays)
        assert hosttest.ht_unresolved() == 0
        ids_c, t_c, _ = _trace_harness(hosttest, h, rays, chaos=True)  # schedule-independent too
    finally:
        hosttest.ht_set_exact(0)
    assert np.array_equal(ids_e, ids_o) and np.array_equal(ids_c, ids_o)
    assert np.array_equal(t_e.view(np.uint32), t_o.view(np.uint32))
    hit = ids_o >= 0
    assert np.array_equal(uv_e[hit].view(np.uint32), uv_o[hit].view(np.uint32))
    hosttest.ht_destroy(h)


def test_exact_tie_mode_render_is_bit_identical(oracle, hosttest, cbox_spheres):
    h = _harness(hosttest, cbox_spheres)
    blocks = _libs.generate_blocks(hosttest, 200, 150, 2, block_size=64)
    acc_o = np.zeros((150, 200, 4), np.float32)
    st = _libs.OrcStats()
    op = _libs.orc_params(max_bounces=8, use_bvh=0, block_size=64)
    assert oracle.orc_render(C.byref(cbox_spheres.view), _libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(acc_o),
                             C.byref(st), 0) == 0
    hosttest.ht_set_exact(1)
    try:
        acc_h = np.zeros_like(acc_o)
        cnt = np.zeros(3, np.uint64)
        hp = _libs.hjk_params(max_bounces=8)
        assert hosttest.ht_render(h, _libs.ptr(blocks), blocks.size, C.byref(hp), _libs.ptr(acc_h), None,
                                  _libs.ptr(cnt)) == 0
        assert hosttest.ht_unresolved() == 0
    finally:
        hosttest.ht_set_exact(0)
    hosttest.ht_destroy(h)
    assert np.array_equal(acc_h.view(np.uint32), acc_o.view(np.uint32))
    assert (int(cnt[1]), int(cnt[2])) == (st.n_extension_rays, st.n_shadow_rays)


@pytest.mark.parametrize("which", ["cbox_spheres", "lattice"])
def test_non_unit_directions_match_the_reference_arithmetic(oracle, hosttest, request, which):
    """The reference's sphere test assumes |d| = 1 and is not geometric otherwise (it accepts an inflated ball at
    a rescaled t); along mirror chains |d| drifts.  The traversal's sphere guard — applied by the nodes flagged
    as having a sphere below them — must still see exactly what the reference arithmetic accepts: directions
    scaled by 1e-3 .. 1e3, first hits compared off ties."""
    scene = request.getfixturevalue("cbox_spheres") if which == "cbox_spheres" else _libs.HostScene.spheres(hosttest, 4)
    h = _harness(hosttest, scene)
    rays = np.concatenate([_libs.camera_rays(scene, 64, 48), _random_rays(scene, 12000, 23)])
    rng = np.random.default_rng(5)
    scale = np.float32(10.0) ** rng.uniform(-3, 3, rays.size).astype(np.float32)
    scale[::7] = np.float32(1.0) + rng.uniform(-1e-6, 1e-6, scale[::7].size).astype(np.float32)  # rounding-level drift
    rays["direction"] *= scale[:, None]
    ids_o, t_o, uv_o, tie = _trace_oracle(oracle, scene, rays, mode=2)
    ids_h, t_h, uv_h = _trace_harness(hosttest, h, rays)
    keep = (tie == 0) & ~np.isinf(t_o)  # "hits at t = +inf" of rays parallel to a triangle's plane: documented deviation
    assert (ids_o[keep] == ids_h[keep]).all(), int((ids_o[keep] != ids_h[keep]).sum())
    hit = keep & (ids_o >= 0)
    assert hit.sum() > 0.3 * rays.size
    assert (t_o[hit].view(np.uint32) == t_h[hit].view(np.uint32)).all()
    sph = hit & (ids_o < scene.info.num_spheres)
    assert sph.sum() > 100  # the sphere test itself was exercised with non-unit directions
    hosttest.ht_destroy(h)


def test_work_statistics_count_the_rays_of_the_render(hosttest, cbox):
    """tools/tree_work.py's counters (node steps / primitive tests per ray of the wide BVH) walk the same rays as the
    render: ray counts equal, every ray takes at least the root step."""
    h = _harness(hosttest, cbox)
    blocks = _libs.generate_blocks(hosttest, 72, 40, 1, block_size=64)
    prm = _libs.hjk_params(max_bounces=4)
    acc = np.zeros((40, 72, 4), np.float32)
    counts = np.zeros(3, np.uint64)
    assert hosttest.ht_render(h, _libs.ptr(blocks), blocks.size, C.byref(prm), _libs.ptr(acc), None, _libs.ptr(counts)) == 0
    hosttest.ht_work_stats.restype = C.c_int
    hosttest.ht_work_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(_abi.HjkParams), C.c_void_p]
    work = np.zeros(6, np.uint64)
    assert hosttest.ht_work_stats(h, _libs.ptr(blocks), blocks.size, C.byref(prm), _libs.ptr(work)) == 0
    assert (work[0], work[3]) == (counts[1], counts[2])
    assert work[1] >= work[0] and work[4] >= work[3] and work[2] > 0
    hosttest.ht_destroy(h)
