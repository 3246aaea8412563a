"""Small render + trace + denoise + GPU BVH build, meant to run under compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hijiki_b200 as hj
ctx = hj.Context(0)
for builder in (0, 1):
    ctx.set_option("bvh_builder", builder)
    for scene in (hj.Scene.from_obj("scenes/cbox/cbox.obj", put_cbox_spheres=True), hj.Scene.spheres(3)):
        ctx.scene_upload(scene.compile())
        w, h = 96, 72
        blocks = hj.ImageBlockGenerator(w, h, 64, 2).blocks()
        for flags in (0, hj.HJK_RENDER_EXACT_TIES):
            ctx.frame_begin(w, h)
            st = ctx.render(blocks, hj.make_params(max_bounces=12, flags=flags))
        img = ctx.readback()
        rays = np.zeros(2000, dtype=hj.RAY_DTYPE)
        rng = np.random.default_rng(1)
        rays["origin"] = rng.standard_normal((2000, 3)).astype(np.float32)
        d = rng.standard_normal((2000, 3)); rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        rays["t_min"], rays["t_max"] = 1e-4, np.inf
        ctx.trace_first_hit(rays); ctx.trace_first_hit(rays, any_hit=True); ctx.trace_first_hit(rays, exact_ties=True)
        print("ok", builder, st.n_rays, float(img[..., :3].mean()), flush=True)
# several tiles per CTA (grid-stride loops, the software pipeline of k_shade) and both trace kernels
ctx.set_option("bvh_builder", 0)
ctx.scene_upload(hj.Scene.from_obj("scenes/cbox/cbox.obj", put_cbox_spheres=True).compile())
tile_blocks = ctx.get_info("blocks_per_sm_tile")
ctx.set_option("blocks_per_sm_tile", 1)
ctx.set_option("blocks_per_sm_traverse", 1)
for coop in (1, 0):
    ctx.set_option("coop_trace", coop)
    blocks = hj.ImageBlockGenerator(320, 200, 64, 3).blocks()
    ctx.frame_begin(320, 200)
    st = ctx.render(blocks, hj.make_params(max_bounces=5))
    print("multi-tile ok", coop, st.n_rays, flush=True)
ctx.set_option("coop_trace", 1)
ctx.set_option("blocks_per_sm_tile", tile_blocks)
ctx.set_option("blocks_per_sm_traverse", 0)
rad = np.ones((72, 96, 4), np.float32); nrm = np.zeros((72, 96, 4), np.float32); nrm[..., 2] = 1
ctx.frame_begin(96, 72)
ctx.denoise_pass(rad, nrm, rad, hj.ImageBlockGenerator(96, 72, 64, 1).blocks(), hj.make_params())
print("denoise ok", float(ctx.readback(normalise=False)[..., 3].mean()))
# k_recon's persistent pipeline with many tiles per CTA (stage reuse, rotation of the warp rectangles), one pass and
# several fused passes, edge tiles on every side (the frame is not a multiple of the tile or of the block)
w, h = 2000, 1100
rng = np.random.default_rng(2)
rad = rng.random((h, w, 4), dtype=np.float32); rad[..., 3] = 1
nrm = rng.standard_normal((h, w, 4)).astype(np.float32)
ctx.frame_begin(w, h)
ctx.denoise_upload(rad, nrm, hj.ImageBlockGenerator(w, h, 128, 1).blocks())
ms = ctx.denoise_resident(hj.make_params(), 2)
print("denoise resident ok", ms, float(ctx.readback(normalise=False)[..., 3].mean()), flush=True)
ctx.set_option("feature_buffers", 1)
blocks = hj.ImageBlockGenerator(640, 360, 128, 6).blocks()
ctx.frame_begin(640, 360)
st = ctx.render(blocks, hj.make_params(max_bounces=3))
print("fused passes ok", st.n_rays, float(ctx.readback()[..., :3].mean()), flush=True)
