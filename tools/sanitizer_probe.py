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
rad = np.ones((72, 96, 4), np.float32); nrm = np.zeros((72, 96, 4), np.float32); nrm[..., 2] = 1
ctx.frame_begin(96, 72)
ctx.denoise_pass(rad, nrm, rad, hj.ImageBlockGenerator(96, 72, 64, 1).blocks(), hj.make_params())
print("denoise ok", float(ctx.readback(normalise=False)[..., 3].mean()))
