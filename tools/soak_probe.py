"""Scratch: long runs of every scene kind through the Python front-end; checks that frames stay finite."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hijiki_b200 as hj
cases = [("cbox+spheres 1080p 512 spp max 1000 bounces", hj.Scene.from_obj("scenes/cbox/cbox.obj", put_cbox_spheres=True), 1920, 1080, 512, 1000),
         ("sphere lattice 4K 32 spp max 64 bounces", hj.Scene.spheres(8), 3840, 2160, 32, 64),
         ("terrain 10M 1080p 128 spp max 8 bounces", hj.Scene.terrain(2237), 1920, 1080, 128, 8)]
for name, scene, w, h, spp, mb in cases:
    t0 = time.time()
    r = hj.Renderer.new(scene, hj.ImageBlockGenerator(w, h, 128, spp), 128, False, max_bounces=mb)
    st = r.render()
    img = r.ctx.readback(normalise=True)
    bad = int((~np.isfinite(img)).any(axis=2).sum())
    print(f"{name}: {st.n_rays / 1e9:.2f} G rays, {st.mrays_per_s:.0f} Mrays/s, {st.n_launches} launches, "
          f"mean {float(np.nanmean(img[..., :3])):.4f}, non-finite texels {bad}, wall {time.time() - t0:.1f} s", flush=True)
    del r
