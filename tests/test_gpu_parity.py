"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs (bit-exact; ties excluded and counted), plus size-independent properties at
BASELINE.json's full sizes.  Needs a B200; nothing here reads /root/reference."""
import ctypes as C

import numpy as np
import pytest

import _libs
import hijiki_b200 as hj
from hijiki_b200 import _abi

pytestmark = pytest.mark.gpu


def _compiled(kind):
    if kind == "cbox":
        return hj.Scene.from_obj(_libs.CBOX_OBJ).compile(use_bvh=True)
    if kind == "cbox_spheres":
        return hj.Scene.from_obj(_libs.CBOX_OBJ, put_cbox_spheres=True).compile(use_bvh=True)
    if kind == "spheres":
        return hj.Scene.spheres(4).compile(use_bvh=True)
    if kind == "terrain":
        return hj.Scene.terrain(64).compile(use_bvh=True)
    raise KeyError(kind)


def _random_rays(compiled, n, seed):
    rng = np.random.default_rng(seed)
    v = compiled.array("vertices")[:, :3]
    lo, hi = v.min(axis=0) - 0.5, v.max(axis=0) + 0.5
    rays = np.zeros(n, dtype=_abi.RAY_DTYPE)
    rays["origin"] = (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["direction"] = d.astype(np.float32)
    rays["t_min"] = 2e-4
    rays["t_max"] = np.inf
    rays["direction"][:6] = np.eye(3, dtype=np.float32).repeat(2, axis=0) * np.array([1, -1] * 3)[:, None]
    return rays


def _oracle_trace(compiled, rays, mode):
    O = _libs.oracle()
    n = rays.size
    ids, t, uv, tie = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros((n, 2), np.float32), np.zeros(n, np.uint8)
    assert O.orc_trace(C.byref(compiled.view), _libs.ptr(rays), n, mode, 1e-4, _libs.ptr(ids), _libs.ptr(t),
                       _libs.ptr(uv), _libs.ptr(tie), 0) == 0
    return ids, t, uv, tie


@pytest.mark.parametrize("kind,mode", [("cbox", 0), ("cbox_spheres", 0), ("spheres", 2), ("terrain", 2)])
def test_first_hit_ids_bit_exact(gpu_ctx, kind, mode):
    """North star: first-hit primitive ids on a fixed ray batch bit-exact vs the reference arithmetic
    (linear scan = the reference default, scene.glsl:134-157); ties excluded and counted."""
    compiled = _compiled(kind)
    gpu_ctx.scene_upload(compiled)
    scene = _libs.HostScene.__new__(_libs.HostScene)
    scene.view, scene.handle, scene.lib = compiled.view, None, None
    rays = np.concatenate([_libs.camera_rays(scene, 160, 90), _random_rays(compiled, 20000, 5)])
    ids_o, t_o, uv_o, tie = _oracle_trace(compiled, rays, mode)
    ids_g, t_g, uv_g = gpu_ctx.trace_first_hit(rays)
    keep = tie == 0
    print(f"{kind}: {rays.size} rays, {int(tie.sum())} ties excluded, "
          f"{int((ids_o[~keep] != ids_g[~keep]).sum())} of them resolved differently")
    assert tie.sum() < 0.002 * rays.size
    assert (ids_o[keep] == ids_g[keep]).all()
    hit = keep & (ids_o >= 0)
    assert hit.sum() > 0.3 * rays.size
    assert (t_o[hit].view(np.uint32) == t_g[hit].view(np.uint32)).all()
    tri = hit & (ids_o >= compiled.info.num_spheres)
    assert (uv_o[tri].view(np.uint32) == uv_g[tri].view(np.uint32)).all()


def test_shadow_rays_exact(gpu_ctx):
    compiled = _compiled("cbox")
    gpu_ctx.scene_upload(compiled)
    rays = _random_rays(compiled, 50000, 9)
    rays["t_max"] = (np.random.default_rng(1).random(rays.size) * 3).astype(np.float32)
    occ = np.zeros(rays.size, np.uint8)
    O = _libs.oracle()
    assert O.orc_occluded(C.byref(compiled.view), _libs.ptr(rays), rays.size, 0, 1e-4, _libs.ptr(occ), 0) == 0
    ids_g, _, _ = gpu_ctx.trace_first_hit(rays, any_hit=True)
    assert (ids_g == occ).all()


def _oracle_render(compiled, blocks, max_bounces, bs, use_bvh):
    O = _libs.oracle()
    w, h = int(blocks[0]["original_dimension"][0]), int(blocks[0]["original_dimension"][1])
    acc = np.zeros((h, w, 4), np.float32)
    st = _libs.OrcStats()
    op = _libs.orc_params(max_bounces=max_bounces, use_bvh=use_bvh, block_size=bs)
    assert O.orc_render(C.byref(compiled.view), _libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(acc),
                        C.byref(st), 0) == 0
    return acc, st


@pytest.mark.parametrize("kind,max_bounces,use_bvh,wave_paths", [
    ("cbox", 8, 0, 4 << 20), ("cbox", 1000, 0, 4 << 20), ("cbox_spheres", 1000, 0, 10000), ("spheres", 16, 2, 4 << 20),
])
def test_render_accumulator_matches_oracle(gpu_ctx, kind, max_bounces, use_bvh, wave_paths):
    """hjk_render vs the oracle's Renderer::render on the same block list: equal ray counts and a
    bit-identical accumulator, tie-affected samples excepted (each touches <= 25 texels)."""
    compiled = _compiled(kind)
    gpu_ctx.scene_upload(compiled)
    gpu_ctx.set_option("wave_paths", wave_paths)  # 10000 forces one pass per wave, several waves
    w, h, bs, spp = 136, 100, 64, 3
    blocks = hj.ImageBlockGenerator(w, h, bs, spp).blocks()
    gpu_ctx.frame_begin(w, h)
    st = gpu_ctx.render(blocks, hj.make_params(max_bounces=max_bounces))
    acc_g = gpu_ctx.readback(normalise=False)
    acc_o, ost = _oracle_render(compiled, blocks, max_bounces, bs, use_bvh)
    gpu_ctx.set_option("wave_paths", 4 << 20)
    diff = (acc_o.view(np.uint32) != acc_g.view(np.uint32)).any(axis=2)
    print(f"{kind}/{max_bounces}: texels differing {int(diff.sum())}/{diff.size}; rays gpu "
          f"{st.n_extension_rays}+{st.n_shadow_rays} oracle {ost.n_extension_rays}+{ost.n_shadow_rays}")
    assert st.n_paths == ost.n_paths == w * h * spp
    assert diff.sum() <= 25 * 4
    # tie-affected paths may take a different number of bounces
    assert abs(st.n_extension_rays - ost.n_extension_rays) <= 100
    assert abs(st.n_shadow_rays - ost.n_shadow_rays) <= 100
    img = gpu_ctx.readback(normalise=True)
    ref = acc_o[..., :3] / acc_o[..., 3:4]
    same = ~diff
    assert np.array_equal(img[..., :3][same], ref[same])  # save_image's divide, src/main.rs:1399


def test_intermediate_layers_match_oracle(gpu_ctx):
    """render.glsl:172-174 outputs of one pass: radiance/1, normal/depth, albedo == 0."""
    compiled = _compiled("cbox")
    gpu_ctx.scene_upload(compiled)
    w, h, bs = 128, 96, 64
    blocks = hj.ImageBlockGenerator(w, h, bs, 1).blocks()
    gpu_ctx.frame_begin(w, h)
    gpu_ctx.render(blocks, hj.make_params(max_bounces=6, flags=hj.HJK_RENDER_NO_RECON))
    O = _libs.oracle()
    layers = np.zeros((3, h, w, 4), np.float32)
    op = _libs.orc_params(max_bounces=6, use_bvh=0, block_size=bs)
    assert O.orc_integrate_frame(C.byref(compiled.view), _libs.ptr(blocks), blocks.size, C.byref(op),
                                 _libs.ptr(layers), None, 0) == 0
    for layer in range(3):
        got = gpu_ctx.read_intermediate(layer)
        bad = (got.view(np.uint32) != layers[layer].view(np.uint32)).any(axis=2).sum()
        assert bad <= 2, (layer, bad)
    assert not gpu_ctx.readback(normalise=False).any()  # NO_RECON leaves the accumulator untouched


def test_denoise_pass_matches_oracle(gpu_ctx):
    w, h, bs = 150, 70, 64
    rng = np.random.default_rng(11)
    blocks = hj.ImageBlockGenerator(w, h, bs, 2).blocks()[:6]
    rad = np.exp(rng.standard_normal((h, w, 4))).astype(np.float32)
    rad[..., 3] = 1.0
    rad[10, 20, 0] = np.nan
    nrm = rng.standard_normal((h, w, 4)).astype(np.float32)
    nrm[..., :3] /= np.linalg.norm(nrm[..., :3], axis=2, keepdims=True)
    alb = rng.random((h, w, 4)).astype(np.float32)
    O = _libs.oracle()
    for albedo in (None, alb):
        gpu_ctx.frame_begin(w, h)
        gpu_ctx.denoise_pass(rad, nrm, albedo, blocks, hj.make_params())
        gpu_ctx.denoise_pass(rad, nrm, albedo, blocks, hj.make_params())  # ADDS (read-modify-write)
        acc_o = np.zeros((h, w, 4), np.float32)
        op = _libs.orc_params(block_size=bs)
        for _ in range(2):
            assert O.orc_reconstruct_frame(_libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(rad), _libs.ptr(nrm),
                                           _libs.ptr(albedo) if albedo is not None else None, _libs.ptr(acc_o), 0) == 0
        assert np.array_equal(acc_o.view(np.uint32), gpu_ctx.readback(normalise=False).view(np.uint32))


@pytest.mark.parametrize("w,h,bs,radius,stddev", [
    (37, 41, 16, 0, 0.5), (150, 70, 64, 1, 0.5), (257, 130, 128, 3, 1.5), (96, 64, 32, 5, 2.0), (130, 75, 64, 8, 3.0),
    (33, 9, 8, 2, 0.5),  # blocks narrower than 2R + 1 texels of interior: every texel takes the general path
])
def test_reconstruction_shapes_match_oracle(gpu_ctx, w, h, bs, radius, stddev):
    """k_recon's interior fast path and its general (block edge / apron) path against the oracle, bit for bit:
    radii 0..8, image sizes that are no multiple of the 32x8 CUDA tile, blocks from 8 to 128 texels."""
    rng = np.random.default_rng(w * 1000 + h)
    # two passes of bs x bs blocks (the reference's generator only makes multiples of 64)
    blocks = np.zeros(2 * ((w + bs - 1) // bs) * ((h + bs - 1) // bs), dtype=_abi.BLOCK_DTYPE)
    k = 0
    for p in range(2):
        so = rng.random(2).astype(np.float32)
        for by in range(0, h, bs):
            for bx in range(0, w, bs):
                blocks[k] = (k, 1000 + k, (bx, by), (min(bs, w - bx), min(bs, h - by)), (w, h), so)
                k += 1
    rad = np.exp(rng.standard_normal((h, w, 4))).astype(np.float32)
    rad[..., 3] = 1.0
    rad[h // 2, w // 3, 1] = np.nan
    nrm = rng.standard_normal((h, w, 4)).astype(np.float32)
    nrm[..., :3] /= np.linalg.norm(nrm[..., :3], axis=2, keepdims=True)
    nrm[: h // 2, : w // 2, :3] = (0.0, 1.0, 0.0)  # a flat region: identical normals, the exp is skipped
    O = _libs.oracle()
    gpu_ctx.frame_begin(w, h)
    acc_o = np.zeros((h, w, 4), np.float32)
    op = _libs.orc_params(block_size=bs, radius=radius, stddev=stddev)
    for one_pass in np.split(blocks, 2):  # a denoise call takes one pass; the second call accumulates
        one_pass = np.ascontiguousarray(one_pass)
        gpu_ctx.denoise_pass(rad, nrm, None, one_pass, hj.make_params(recon_radius=radius, recon_stddev=stddev))
        assert O.orc_reconstruct_frame(_libs.ptr(one_pass), one_pass.size, C.byref(op), _libs.ptr(rad), _libs.ptr(nrm),
                                       None, _libs.ptr(acc_o), 0) == 0
    assert np.array_equal(acc_o.view(np.uint32), gpu_ctx.readback(normalise=False).view(np.uint32))


def test_device_math_matches_spec_bitwise(gpu_ctx):
    """The device build of hjk_math.cuh == oracle/orc_math.h, through the only place it surfaces
    directly: the reconstruction weights (exp) — checked via a one-texel splat."""
    O = _libs.oracle()
    w = h = 64
    for so in [(0.0, 0.0), (0.5, 0.5), (0.123, 0.877), (0.999, 0.001)]:
        blocks = hj.ImageBlockGenerator(w, h, 64, 1).blocks()
        blocks["sample_offset"] = so
        rad = np.zeros((h, w, 4), np.float32)
        rad[32, 32] = (1, 1, 1, 1)
        nrm = np.zeros((h, w, 4), np.float32)
        gpu_ctx.frame_begin(w, h)
        gpu_ctx.denoise_pass(rad, nrm, None, blocks, hj.make_params())
        acc = gpu_ctx.readback(normalise=False)
        wts = np.zeros(25, np.float32)
        O.orc_recon_spatial_weights(2, 0.5, so[0], so[1], _libs.ptr(wts))
        # output texel (32-dx, 32-dy) receives tap (dx, dy)
        for dx in range(-2, 3):
            for dy in range(-2, 3):
                wt = wts[(dx + 2) * 5 + (dy + 2)]
                expect = np.float32(0.0) if wt < 0 else wt
                assert acc[32 - dy, 32 - dx, 3].view(np.uint32) == np.float32(expect).view(np.uint32)


def test_render_is_deterministic_and_wave_invariant(gpu_ctx):
    """Same block list -> same bits, whatever the wave size (queue order differs, results do not)."""
    compiled = _compiled("cbox_spheres")
    gpu_ctx.scene_upload(compiled)
    w, h = 200, 120
    blocks = hj.ImageBlockGenerator(w, h, 64, 4).blocks()
    outs = []
    for wave in (4 << 20, 30000, 4 << 20):
        gpu_ctx.set_option("wave_paths", wave)
        gpu_ctx.frame_begin(w, h)
        gpu_ctx.render(blocks, hj.make_params(max_bounces=12))
        outs.append(gpu_ctx.readback(normalise=False))
    gpu_ctx.set_option("wave_paths", 4 << 20)
    assert np.array_equal(outs[0], outs[2])
    # different wave sizes change how many passes one reconstruction launch folds, not the add order
    assert np.array_equal(outs[0], outs[1])


def test_sample_pass_split_is_linear(gpu_ctx):
    """rendering passes p = r mod N on N contexts and summing the accumulators equals the
    single-context frame up to fp32 summation order."""
    compiled = _compiled("cbox")
    gpu_ctx.scene_upload(compiled)
    w, h = 160, 96
    gen = hj.ImageBlockGenerator(w, h, 64, 4)
    blocks = gen.blocks()
    gpu_ctx.frame_begin(w, h)
    gpu_ctx.render(blocks, hj.make_params(max_bounces=8))
    full = gpu_ctx.readback(normalise=False)
    total = np.zeros_like(full, dtype=np.float64)
    for r in range(2):
        gpu_ctx.frame_begin(w, h)
        gpu_ctx.render(hj.split_passes(blocks, gen.blocks_per_pass, r, 2), hj.make_params(max_bounces=8))
        total += gpu_ctx.readback(normalise=False)
    assert np.allclose(total, full, rtol=2e-6, atol=1e-6)


def test_full_size_pass_properties(gpu_ctx):
    """BASELINE config 2 geometry (cbox 1920x1080, max 8 bounces), two passes: every texel gets weight,
    radiance is finite and non-negative, per-texel weight stays below the analytic tap-sum bound."""
    compiled = _compiled("cbox")
    gpu_ctx.scene_upload(compiled)
    w, h = 1920, 1080
    blocks = hj.ImageBlockGenerator(w, h, 128, 2).blocks()
    gpu_ctx.frame_begin(w, h)
    st = gpu_ctx.render(blocks, hj.make_params(max_bounces=8))
    acc = gpu_ctx.readback(normalise=False)
    assert st.n_paths == 2 * w * h
    assert 2 * w * h <= st.n_extension_rays <= 8 * 2 * w * h
    assert 0 < st.n_shadow_rays <= st.n_extension_rays
    assert np.isfinite(acc).all() and (acc >= 0).all()
    assert (acc[..., 3] > 0).all()
    assert acc[..., 3].max() <= 2 * 1.7
    img = acc[..., :3] / acc[..., 3:4]
    assert 0.05 < img.mean() < 1.0
    print(f"1080p x2 passes: {st.n_rays / 1e6:.1f} Mrays in {st.ms_total:.2f} ms = {st.mrays_per_s:.0f} Mrays/s")


def test_error_paths_return_codes(gpu_ctx):
    lib = gpu_ctx.lib
    fresh = hj.Context(0)
    with pytest.raises(hj.HijikiError) as e:
        fresh.render(hj.ImageBlockGenerator(64, 64, 64, 1).blocks(), hj.make_params())
    assert e.value.status == -3  # HJK_ERR_NO_SCENE
    with pytest.raises(hj.HijikiError) as e:
        fresh.readback()
    assert e.value.status == -4
    fresh.close()
    bad = hj.ImageBlockGenerator(64, 64, 64, 1).blocks()
    bad["origin"][0] = (32, 0)  # sticks out of the image
    compiled = _compiled("cbox")
    gpu_ctx.scene_upload(compiled)
    with pytest.raises(hj.HijikiError) as e:
        gpu_ctx.render(bad, hj.make_params())
    assert e.value.status == -1
    assert lib.hjk_destroy(None) == 0


def test_exact_tie_mode_matches_reference_on_every_ray(gpu_ctx):
    """HJK_RENDER_EXACT_TIES / trace mode bit 1: the linear-scan winner among hits closer than M_EPS is
    reproduced, so nothing has to be excluded: ids, t, uv bit-exact on all rays, and the frame
    accumulator bit-identical to the oracle's."""
    from test_pipeline_host import _tie_heavy_rays
    compiled = _compiled("cbox")
    gpu_ctx.scene_upload(compiled)
    rays = _tie_heavy_rays()
    ids_o, t_o, uv_o, tie = _oracle_trace(compiled, rays, 0)
    assert tie.sum() >= 10
    ids_d, _, _ = gpu_ctx.trace_first_hit(rays)
    assert ((ids_d != ids_o) & (tie == 0)).sum() == 0
    ids_e, t_e, uv_e = gpu_ctx.trace_first_hit(rays, exact_ties=True)
    assert gpu_ctx.get_info("unresolved_ties") == 0
    print(f"ties {int(tie.sum())}: default mode resolves {int((ids_d != ids_o).sum())} differently, exact mode "
          f"{int((ids_e != ids_o).sum())}")
    assert np.array_equal(ids_e, ids_o)
    assert np.array_equal(t_e.view(np.uint32), t_o.view(np.uint32))
    hit = ids_o >= 0
    assert np.array_equal(uv_e[hit].view(np.uint32), uv_o[hit].view(np.uint32))
    # whole frames, three scenes
    for kind, use_bvh, bounces in (("cbox", 0, 8), ("cbox_spheres", 0, 1000), ("spheres", 2, 8)):
        compiled = _compiled(kind)
        gpu_ctx.scene_upload(compiled)
        w, h, bs, spp = 200, 150, 64, 2
        blocks = hj.ImageBlockGenerator(w, h, bs, spp).blocks()
        gpu_ctx.frame_begin(w, h)
        st = gpu_ctx.render(blocks, hj.make_params(max_bounces=bounces, flags=hj.HJK_RENDER_EXACT_TIES))
        acc_g = gpu_ctx.readback(normalise=False)
        acc_o, ost = _oracle_render(compiled, blocks, bounces, bs, use_bvh)
        unresolved = gpu_ctx.get_info("unresolved_ties")
        diff = (acc_o.view(np.uint32) != acc_g.view(np.uint32)).any(axis=2)
        print(f"{kind}: exact-tie frame, texels differing {int(diff.sum())}, unresolved clusters {unresolved}")
        if unresolved == 0:
            assert diff.sum() == 0
            assert (st.n_extension_rays, st.n_shadow_rays) == (ost.n_extension_rays, ost.n_shadow_rays)
        else:
            assert diff.sum() <= 25 * unresolved


def test_async_render_then_synchronize(gpu_ctx):
    """HJK_RENDER_ASYNC returns after enqueue (the reference's render() never waits for the device either,
    src/main.rs:1487-1490); synchronize + readback then see the finished frame."""
    compiled = _compiled("cbox")
    gpu_ctx.scene_upload(compiled)
    w, h = 160, 96
    blocks = hj.ImageBlockGenerator(w, h, 64, 2).blocks()
    gpu_ctx.frame_begin(w, h)
    gpu_ctx.render(blocks, hj.make_params(max_bounces=6))
    want = gpu_ctx.readback(normalise=False)
    gpu_ctx.frame_begin(w, h)
    assert gpu_ctx.render(blocks, hj.make_params(max_bounces=6, flags=hj.HJK_RENDER_ASYNC), want_stats=False) is None
    gpu_ctx.synchronize()
    assert np.array_equal(gpu_ctx.readback(normalise=False), want)
    # two renders into the same frame accumulate (src/main.rs:1330: the reconstruction ADDS)
    gpu_ctx.render(blocks, hj.make_params(max_bounces=6), want_stats=False)
    twice = gpu_ctx.readback(normalise=False)
    assert np.allclose(twice, 2 * want, rtol=1e-6)


def test_converged_image_matches_reference_statistically(gpu_ctx):
    """North-star image bar: with DIFFERENT block seeds and sub-pixel offsets (so no sample is shared) the
    CUDA image and the oracle image estimate the same radiance: mean relative error of the frame mean
    < 1 %, and the per-pixel RMSE relative to the mean radiance is what two independent 256-spp estimates
    of this scene give each other (measured between two oracle runs) within 25 %."""
    compiled = _compiled("cbox")
    gpu_ctx.scene_upload(compiled)
    w, h, bs, spp, bounces = 96, 72, 64, 256, 8
    O = _libs.oracle()

    def oracle_image(seed):
        blocks = hj.ImageBlockGenerator(w, h, bs, spp, root_seed=seed).blocks()
        acc = np.zeros((h, w, 4), np.float32)
        op = _libs.orc_params(max_bounces=bounces, use_bvh=1, block_size=bs)
        assert O.orc_render(C.byref(compiled.view), _libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(acc), None,
                            0) == 0
        return acc[..., :3] / acc[..., 3:4]

    ref_a, ref_b = oracle_image(11), oracle_image(22)
    gpu_ctx.frame_begin(w, h)
    gpu_ctx.render(hj.ImageBlockGenerator(w, h, bs, spp, root_seed=33).blocks(), hj.make_params(max_bounces=bounces))
    img = gpu_ctx.readback(normalise=True)[..., :3]
    mean = float(ref_a.mean())
    rmse_refs = float(np.sqrt(np.mean((ref_a - ref_b) ** 2))) / mean
    rmse_gpu = float(np.sqrt(np.mean((img - ref_a) ** 2))) / mean
    mre = abs(float(img.mean()) - mean) / mean
    print(f"relative RMSE oracle/oracle {rmse_refs:.4f}, cuda/oracle {rmse_gpu:.4f}; mean relative error {mre:.5f}")
    assert mre < 0.01
    assert rmse_gpu < 1.25 * rmse_refs + 1e-3


@pytest.mark.parametrize("kind,mode", [("cbox", 0), ("cbox_spheres", 0), ("spheres", 2), ("terrain", 2)])
def test_gpu_built_bvh_gives_the_same_hits(gpu_ctx, kind, mode):
    """the wide BVH built ON the GPU (LBVH + collapse, bvh_build_gpu.cuh) passes the host
    structural check, and — hits not depending on the tree — first-hit ids/t stay bit-exact off ties
    and whole frames bit-identical in exact-tie mode."""
    compiled = _compiled(kind)
    gpu_ctx.set_option("bvh_builder", 1)
    gpu_ctx.set_option("bvh_validate", 1)
    try:
        gpu_ctx.scene_upload(compiled)  # fails if the structural check fails
        assert gpu_ctx.get_info("bvh_build_us") > 0, "the GPU builder was not used"
        scene = _libs.HostScene.__new__(_libs.HostScene)
        scene.view, scene.handle, scene.lib = compiled.view, None, None
        rays = np.concatenate([_libs.camera_rays(scene, 120, 80), _random_rays(compiled, 15000, 6)])
        ids_o, t_o, uv_o, tie = _oracle_trace(compiled, rays, mode)
        ids_g, t_g, _ = gpu_ctx.trace_first_hit(rays)
        keep = tie == 0
        assert (ids_o[keep] == ids_g[keep]).all()
        hit = keep & (ids_o >= 0)
        assert (t_o[hit].view(np.uint32) == t_g[hit].view(np.uint32)).all()
        ids_e, t_e, _ = gpu_ctx.trace_first_hit(rays, exact_ties=True)
        if gpu_ctx.get_info("unresolved_ties") == 0:
            # axis-parallel probe rays that escape the scene are "hit at t = +inf" by the reference
            # arithmetic (documented deviation): excluded
            real = ~np.isinf(t_o)
            assert np.array_equal(ids_e[real], ids_o[real])
        w, h, bs, spp, bounces = 136, 100, 64, 2, 8
        blocks = hj.ImageBlockGenerator(w, h, bs, spp).blocks()
        gpu_ctx.frame_begin(w, h)
        gpu_ctx.render(blocks, hj.make_params(max_bounces=bounces, flags=hj.HJK_RENDER_EXACT_TIES))
        acc_g = gpu_ctx.readback(normalise=False)
        acc_o, _ = _oracle_render(compiled, blocks, bounces, bs, mode)
        diff = (acc_o.view(np.uint32) != acc_g.view(np.uint32)).any(axis=2)
        assert diff.sum() <= 25 * gpu_ctx.get_info("unresolved_ties")
        print(f"{kind}: GPU build {gpu_ctx.get_info('bvh_build_us')} us, {gpu_ctx.get_info('bvh_nodes')} nodes, depth "
              f"{gpu_ctx.get_info('bvh_depth')}")
    finally:
        gpu_ctx.set_option("bvh_builder", 0)
        gpu_ctx.set_option("bvh_validate", 0)


@pytest.mark.parametrize("coop,batch_cost", [(1, 180), (1, 0), (0, 180)])
@pytest.mark.parametrize("kind,max_bounces,use_bvh", [("cbox", 8, 0), ("cbox_spheres", 1000, 0), ("spheres", 16, 2)])
def test_trace_kernel_variants_match_oracle(gpu_ctx, kind, max_bounces, use_bvh, coop, batch_cost):
    """The default-mode trace kernel comes in two forms: k_trace_coop (default; a warp pools its
    primitive tests and spreads them over all 32 lanes when that is cheaper, batch_cost 0 = always)
    and the per-lane k_trace (coop_trace=0).  Same hits off ties, same frames within the tie budget."""
    compiled = _compiled(kind)
    gpu_ctx.scene_upload(compiled)
    gpu_ctx.set_option("coop_trace", coop)
    gpu_ctx.set_option("coop_batch_cost", batch_cost)
    try:
        w, h, bs, spp = 136, 100, 64, 3
        blocks = hj.ImageBlockGenerator(w, h, bs, spp).blocks()
        gpu_ctx.frame_begin(w, h)
        st = gpu_ctx.render(blocks, hj.make_params(max_bounces=max_bounces))
        acc_g = gpu_ctx.readback(normalise=False)
    finally:
        gpu_ctx.set_option("coop_trace", 1)
        gpu_ctx.set_option("coop_batch_cost", -1)
    acc_o, ost = _oracle_render(compiled, blocks, max_bounces, bs, use_bvh)
    diff = (acc_o.view(np.uint32) != acc_g.view(np.uint32)).any(axis=2)
    print(f"coop={coop}/{batch_cost} {kind}: texels differing {int(diff.sum())}; rays {st.n_extension_rays}+{st.n_shadow_rays} vs "
          f"{ost.n_extension_rays}+{ost.n_shadow_rays}")
    assert st.n_paths == ost.n_paths
    assert diff.sum() <= 25 * 4
    assert abs(st.n_extension_rays - ost.n_extension_rays) <= 100


@pytest.mark.parametrize("w,h,spp,max_bounces,rr_start", [
    (1, 1, 3, 8, 3),        # a single texel
    (7, 5, 2, 2, 3),        # a frame smaller than the 32x8 reconstruction tile
    (40, 33, 2, 1, 3),      # one bounce: next-event estimation once
    (65, 64, 2, 12, 0),     # roulette from the first bounce; one texel column spills into a second block
    (64, 130, 1, 6, 100),   # roulette never starts within the bounce limit
])
def test_render_edge_shapes_match_oracle(gpu_ctx, w, h, spp, max_bounces, rr_start):
    """Frame and parameter corner cases through hjk_render against the oracle (exact-tie mode: bit-identical)."""
    compiled = _compiled("cbox_spheres")
    gpu_ctx.scene_upload(compiled)
    bs = 64
    blocks = hj.ImageBlockGenerator(w, h, bs, spp).blocks()
    gpu_ctx.frame_begin(w, h)
    st = gpu_ctx.render(blocks, hj.make_params(max_bounces=max_bounces, rr_start=rr_start, flags=hj.HJK_RENDER_EXACT_TIES))
    acc_g = gpu_ctx.readback(normalise=False)
    O = _libs.oracle()
    acc_o = np.zeros((h, w, 4), np.float32)
    ost = _libs.OrcStats()
    op = _libs.orc_params(max_bounces=max_bounces, rr_start=rr_start, use_bvh=0, block_size=bs)
    assert O.orc_render(C.byref(compiled.view), _libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(acc_o),
                        C.byref(ost), 0) == 0
    assert (st.n_paths, st.n_extension_rays, st.n_shadow_rays) == (ost.n_paths, ost.n_extension_rays, ost.n_shadow_rays)
    diff = (acc_o.view(np.uint32) != acc_g.view(np.uint32)).any(axis=2)
    assert diff.sum() <= 25 * gpu_ctx.get_info("unresolved_ties")


@pytest.mark.parametrize("kind,max_bounces", [("cbox_spheres", 12), ("spheres", 16)])
def test_default_mode_is_order_independent(gpu_ctx, kind, max_bounces):
    """The default closest-hit rule (smallest t, equal t by the lower shape id) is a function of the hit set:
    pooled or per-lane primitive tests, one trace kernel or the other, small or large waves — same frame, bit
    for bit, ties included."""
    compiled = _compiled(kind)
    gpu_ctx.scene_upload(compiled)
    w, h, bs, spp = 136, 100, 64, 4
    blocks = hj.ImageBlockGenerator(w, h, bs, spp).blocks()
    frames, counts = [], []
    try:
        for coop, cost, wave in [(1, 180, 64 << 20), (1, 0, 64 << 20), (0, 180, 64 << 20), (1, 180, 20000), (1, 400, 64 << 20)]:
            gpu_ctx.set_option("coop_trace", coop)
            gpu_ctx.set_option("coop_batch_cost", cost)
            gpu_ctx.set_option("wave_paths", wave)
            gpu_ctx.frame_begin(w, h)
            st = gpu_ctx.render(blocks, hj.make_params(max_bounces=max_bounces))
            frames.append(gpu_ctx.readback(normalise=False))
            counts.append((st.n_paths, st.n_extension_rays, st.n_shadow_rays))
    finally:
        gpu_ctx.set_option("coop_trace", 1)
        gpu_ctx.set_option("coop_batch_cost", -1)
        gpu_ctx.set_option("wave_paths", 64 << 20)
    assert len(set(counts)) == 1, counts
    for f in frames[1:]:
        assert np.array_equal(frames[0].view(np.uint32), f.view(np.uint32))


@pytest.mark.parametrize("kind,builder", [("cbox_spheres", 0), ("spheres", 0), ("cbox_spheres", 1)])
def test_non_unit_directions_match_the_reference_arithmetic(gpu_ctx, kind, builder):
    """Sphere guard on the device (per-node 'sphere below' flag from either builder): directions scaled by 1e-3 .. 1e3 — where the reference's sphere test is no longer geometric — still
    give the reference arithmetic's first hits, off ties."""
    compiled = _compiled(kind)
    gpu_ctx.set_option("bvh_builder", builder)
    try:
        gpu_ctx.scene_upload(compiled)
    finally:
        gpu_ctx.set_option("bvh_builder", 0)
    assert gpu_ctx.get_info("sphere_guard") == (2 if kind == "spheres" else 1)  # both builders flag nodes per subtree
    scene = _libs.HostScene.__new__(_libs.HostScene)
    scene.view, scene.handle, scene.lib = compiled.view, None, None
    rays = np.concatenate([_libs.camera_rays(scene, 64, 48), _random_rays(compiled, 12000, 23)])
    rng = np.random.default_rng(5)
    scale = np.float32(10.0) ** rng.uniform(-3, 3, rays.size).astype(np.float32)
    scale[::7] = np.float32(1.0) + rng.uniform(-1e-6, 1e-6, scale[::7].size).astype(np.float32)
    rays["direction"] *= scale[:, None]
    ids_o, t_o, _, tie = _oracle_trace(compiled, rays, 2)
    ids_g, t_g, _ = gpu_ctx.trace_first_hit(rays)
    keep = (tie == 0) & ~np.isinf(t_o)
    assert (ids_o[keep] == ids_g[keep]).all(), int((ids_o[keep] != ids_g[keep]).sum())
    hit = keep & (ids_o >= 0)
    assert (t_o[hit].view(np.uint32) == t_g[hit].view(np.uint32)).all()
    assert (hit & (ids_o < compiled.info.num_spheres)).sum() > 100


def test_scene_without_emitters_is_refused(gpu_ctx):
    """sampleEmitter (scene.glsl:54-89) reads emitters[0] at every diffuse hit: a scene without an emissive shape
    has no defined result in the reference; the upload refuses it instead of reading unwritten memory."""
    D = _abi.MAT_DIFFUSE
    verts = [(-1, 0, 0, 0, 0, 0, 1, 0), (1, 0, 0, 1, 0, 0, 1, 0), (0, 1, 0, 0, 0, 0, 1, 1)]
    half = 0.0
    scene = _libs.CustomScene(((0.0, 0.3, 4.0), (np.sin(half), 0.0, 0.0, np.cos(half)), 40.0), triangles=[(0, 1, 2)],
                              vertices=verts, materials=[(D, 0)], diffuse=[(0.5, 0.5, 0.5, 0)])
    assert scene.info.num_emitters == 0
    with pytest.raises(hj.HijikiError) as e:
        gpu_ctx.scene_upload(scene)
    assert e.value.status == -1 and "emitter" in str(e.value)
    gpu_ctx.scene_upload(_compiled("cbox"))  # the context is still usable


@pytest.mark.parametrize("threshold", [0, 1, 32])
def test_fetch_threshold_extremes_terminate_and_match(gpu_ctx, threshold):
    """Option fetch_threshold over its whole accepted range (0 used to leave an idle warp without work for ever):
    same frame as the default, bit for bit."""
    compiled = _compiled("cbox_spheres")
    gpu_ctx.scene_upload(compiled)
    w, h = 136, 100
    blocks = hj.ImageBlockGenerator(w, h, 64, 2).blocks()
    gpu_ctx.frame_begin(w, h)
    gpu_ctx.render(blocks, hj.make_params(max_bounces=12))
    want = gpu_ctx.readback(normalise=False)
    gpu_ctx.set_option("fetch_threshold", threshold)
    try:
        gpu_ctx.frame_begin(w, h)
        gpu_ctx.render(blocks, hj.make_params(max_bounces=12))
        got = gpu_ctx.readback(normalise=False)
    finally:
        gpu_ctx.set_option("fetch_threshold", -1)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
    assert gpu_ctx.get_info("stack_overflows") == 0


def test_feature_buffers_average_the_first_hit_features(gpu_ctx):
    """Option feature_buffers: hjk_read_features returns, per texel, the mean over its samples of layer 1 =
    (normal, depth) (render.glsl:173) — checked against the oracle's per-pass layers."""
    compiled = _compiled("cbox")
    gpu_ctx.scene_upload(compiled)
    w, h, bs, spp = 128, 96, 64, 3
    gen = hj.ImageBlockGenerator(w, h, bs, spp)
    blocks = gen.blocks()
    gpu_ctx.set_option("feature_buffers", 1)
    try:
        gpu_ctx.frame_begin(w, h)
        gpu_ctx.render(blocks, hj.make_params(max_bounces=4))
        feat = gpu_ctx.read_features()
        acc = gpu_ctx.readback(normalise=False)
    finally:
        gpu_ctx.set_option("feature_buffers", 0)
    O = _libs.oracle()
    op = _libs.orc_params(max_bounces=4, use_bvh=0, block_size=bs)
    total = np.zeros((h, w, 4), np.float32)
    for p in range(spp):
        one = np.ascontiguousarray(blocks[p * gen.blocks_per_pass:(p + 1) * gen.blocks_per_pass])
        layers = np.zeros((3, h, w, 4), np.float32)
        assert O.orc_integrate_frame(C.byref(compiled.view), _libs.ptr(one), one.size, C.byref(op), _libs.ptr(layers),
                                     None, 0) == 0
        total = total + layers[1]
    want = total / np.float32(spp)
    bad = (want.view(np.uint32) != feat.view(np.uint32)).any(axis=2).sum()
    print(f"feature buffers: texels differing from the oracle's mean {int(bad)}/{w * h}")
    assert bad <= 6  # a tie at the first hit can move one texel's normal
    # the accumulator is the same frame with or without the feature sums
    gpu_ctx.frame_begin(w, h)
    gpu_ctx.render(blocks, hj.make_params(max_bounces=4))
    assert np.array_equal(acc.view(np.uint32), gpu_ctx.readback(normalise=False).view(np.uint32))
    with pytest.raises(hj.HijikiError):
        gpu_ctx.read_features()  # option off


def _chain_scene(n=700, ratio=1.12):
    """Triangles shrinking geometrically towards the origin: the SAH builder peels them off scale by scale and the
    wide BVH comes out 15 levels deep — twice the depth of any BASELINE scene, deeper than the eight stack entries the
    trace kernels used to keep in shared memory."""
    verts, tris, mats = [], [], []
    for k in range(n):
        s = ratio ** (-k)
        z = 0.2 * s  # no two in one plane, no two overlapping: hits of different triangles differ in t
        verts += [(s, 0, z, 0, 0, 0, 1, 0), (1.1 * s, 0, z, 1, 0, 0, 1, 0), (s, 0.3 * s, z, 0, 0, 0, 1, 1)]
        tris.append((3 * k, 3 * k + 1, 3 * k + 2))
        mats.append((_abi.MAT_EMISSIVE if k % 50 == 0 else _abi.MAT_DIFFUSE, 0))
    return _libs.CustomScene(((0.4, 0.1, 2.5), (0.0, 0.0, 0.0, 1.0), 35.0), triangles=tris, vertices=verts,
                             materials=mats, diffuse=[(0.6, 0.6, 0.6, 0)], emissive=[(5, 5, 5, 0)])


@pytest.mark.parametrize("n_tris,min_depth", [(700, 14), (300, 20)])
def test_deep_tree_traversal_stack(gpu_ctx, n_tris, min_depth):
    """Wide BVHs of 14 and 23 levels (the exact SAH sweep peels the 300-triangle chain almost one per level; the
    700-triangle one would come out 82 deep and is rebuilt with binned splits): the launch sizes the shared-memory
    traversal stack to the tree (depth + 1 entries, twice that where primitive groups are postponed and it fits);
    hits stay exact and no stack entry is ever dropped."""
    scene = _chain_scene(n_tris)
    gpu_ctx.scene_upload(scene)
    depth = gpu_ctx.get_info("bvh_depth")
    assert min_depth <= depth <= 32, depth
    rng = np.random.default_rng(12)
    n = 30000
    k = rng.integers(0, min(160, n_tris), n)  # down to triangles of 1e-8: the deepest levels of the tree
    s = 1.12 ** (-k.astype(np.float64))
    target = np.stack([s * (1.0 + 0.1 * rng.random(n)), 0.3 * s * rng.random(n), 0.2 * s], axis=1)
    origin = np.stack([rng.uniform(-0.2, 1.5, n), rng.uniform(-0.3, 0.6, n), rng.uniform(0.5, 2.0, n)], axis=1)
    d = target - origin
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(n, dtype=_abi.RAY_DTYPE)
    rays["origin"], rays["direction"] = origin.astype(np.float32), d.astype(np.float32)
    rays["t_min"], rays["t_max"] = 1e-4, np.inf
    eps = 1e-9  # M_EPS as a parameter: at 1e-4 every triangle below that size would be one tie cluster
    O = _libs.oracle()
    ids_o, t_o, uv_o, tie = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros((n, 2), np.float32), np.zeros(n, np.uint8)
    assert O.orc_trace(C.byref(scene.view), _libs.ptr(rays), n, 2, eps, _libs.ptr(ids_o), _libs.ptr(t_o), _libs.ptr(uv_o),
                       _libs.ptr(tie), 0) == 0
    ids_g, t_g, uv_g = gpu_ctx.trace_first_hit(rays, eps=eps)
    keep = (tie == 0) & np.isfinite(t_o)
    assert (ids_o >= 0).mean() > 0.2 and keep.mean() > 0.85
    ids_e, _, _ = gpu_ctx.trace_first_hit(rays, exact_ties=True, eps=eps)
    assert (ids_e != ids_o).sum() <= gpu_ctx.get_info("unresolved_ties")
    assert np.array_equal(ids_o[keep], ids_g[keep])
    hit = keep & (ids_o >= 0)
    assert np.array_equal(t_o[hit].view(np.uint32), t_g[hit].view(np.uint32))
    # whole frames through both trace kernels (default = pooled k_trace_coop, exact-tie = per-lane k_trace)
    w, h, bs = 96, 64, 64
    blocks = hj.ImageBlockGenerator(w, h, bs, 2).blocks()
    acc_o, ost = _oracle_render(scene, blocks, 8, bs, 2)
    for flags in (0, hj.HJK_RENDER_EXACT_TIES):
        gpu_ctx.frame_begin(w, h)
        st = gpu_ctx.render(blocks, hj.make_params(max_bounces=8, flags=flags))
        acc_g = gpu_ctx.readback(normalise=False)
        diff = (acc_o.view(np.uint32) != acc_g.view(np.uint32)).any(axis=2)
        budget = 25 * (4 if flags == 0 else gpu_ctx.get_info("unresolved_ties"))
        assert diff.sum() <= budget, (flags, int(diff.sum()))
    assert gpu_ctx.get_info("stack_overflows") == 0
    print(f"chain scene: wide BVH depth {depth}, {int((ids_o >= 0).sum())} of {n} probe rays hit, {int(tie.sum())} ties")


@pytest.mark.parametrize("kind", ["cbox", "cbox_spheres", "spheres"])
def test_shade_with_and_without_the_material_sort_agree(gpu_ctx, kind):
    """k_shade<SORT>: sorting a tile's hits by material only changes which thread shades which hit."""
    compiled = _compiled(kind)
    gpu_ctx.scene_upload(compiled)
    assert gpu_ctx.get_info("shade_sort") == (0 if kind == "cbox" else 1)  # the per-scene default
    w, h = 136, 100
    blocks = hj.ImageBlockGenerator(w, h, 64, 3).blocks()
    frames = []
    try:
        for mode in (0, 1):
            gpu_ctx.set_option("shade_sort", mode)
            gpu_ctx.frame_begin(w, h)
            st = gpu_ctx.render(blocks, hj.make_params(max_bounces=12))
            frames.append((gpu_ctx.readback(normalise=False), st.n_rays))
    finally:
        gpu_ctx.set_option("shade_sort", -1)
    assert frames[0][1] == frames[1][1]
    assert np.array_equal(frames[0][0].view(np.uint32), frames[1][0].view(np.uint32))


def test_asynchronous_readback(gpu_ctx):
    """hjk_readback_begin / hjk_readback_wait: the staged frame survives the next hjk_frame_begin + hjk_render."""
    import torch
    compiled = _compiled("cbox")
    gpu_ctx.scene_upload(compiled)
    w, h = 200, 120
    gen = hj.ImageBlockGenerator(w, h, 64, 2)
    blocks = gen.blocks()
    first, second = blocks[:gen.blocks_per_pass], blocks[gen.blocks_per_pass:]
    want = []
    for part in (first, second):
        gpu_ctx.frame_begin(w, h)
        gpu_ctx.render(np.ascontiguousarray(part), hj.make_params(max_bounces=6))
        want.append(gpu_ctx.readback(normalise=True))
    host = [torch.empty((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    for k, part in enumerate((first, second)):
        gpu_ctx.frame_begin(w, h)
        gpu_ctx.render(np.ascontiguousarray(part), hj.make_params(max_bounces=6), want_stats=False)
        gpu_ctx.readback_begin_ptr(-1, host[k].data_ptr(), w * 16, normalise=True)  # returns before the copy is done
    gpu_ctx.readback_wait()
    for k in range(2):
        assert np.array_equal(host[k].numpy().view(np.uint32), want[k].view(np.uint32))


def test_reconstruction_pipeline_many_tiles_per_cta(gpu_ctx):
    """k_recon's persistent pipeline with a dozen tiles per CTA (stage reuse, rectangles rotating over the warps), a frame
    that is a multiple of neither the tile nor the block: five passes folded by one launch (the accumulator texel stays
    in a register across them) against one launch per pass (it travels through the stage every time) — same bits, in
    the accumulator and in the feature sums."""
    compiled = _compiled("cbox")
    gpu_ctx.scene_upload(compiled)
    w, h, spp = 1610, 907, 5
    blocks = hj.ImageBlockGenerator(w, h, 128, spp).blocks()
    outs = []
    gpu_ctx.set_option("feature_buffers", 1)
    try:
        for wave in (64 << 20, w * h):
            gpu_ctx.set_option("wave_paths", wave)
            gpu_ctx.frame_begin(w, h)
            gpu_ctx.render(blocks, hj.make_params(max_bounces=2))
            outs.append((gpu_ctx.readback(normalise=False), gpu_ctx.read_features()))
    finally:
        gpu_ctx.set_option("feature_buffers", 0)
        gpu_ctx.set_option("wave_paths", 4 << 20)
    assert np.isfinite(outs[0][0]).all() and outs[0][0][..., 3].min() > 0
    assert np.array_equal(outs[0][0].view(np.uint32), outs[1][0].view(np.uint32))
    assert np.array_equal(outs[0][1].view(np.uint32), outs[1][1].view(np.uint32))
