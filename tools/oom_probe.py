"""Scratch: a wave that cannot fit in HBM is halved until it does (hjk_render's allocation fallback)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hijiki_b200 as hj
W, H, spp = 1920, 1080, 512
ctx = hj.Context(0); ctx.scene_upload(hj.Scene.from_obj('scenes/cbox/cbox.obj').compile())
blocks = hj.ImageBlockGenerator(W, H, 128, spp).blocks()
p = hj.make_params(max_bounces=8)
ctx.frame_begin(W, H); st = ctx.render(blocks, p); ref = ctx.readback(normalise=False)
print(f"default wave: {st.mrays_per_s:.0f} Mrays/s, {st.n_launches} launches")
ctx.set_option("wave_paths", 2**31 - 1)   # 512 passes of 1080p = 1.06 G slots = 212 GB: does not fit
ctx.frame_begin(W, H); st = ctx.render(blocks, p); got = ctx.readback(normalise=False)
diff = int((ref.view(np.uint32) != got.view(np.uint32)).any(axis=2).sum())
print(f"oversized wave: {st.mrays_per_s:.0f} Mrays/s, {st.n_launches} launches, texels differing from the default "
      f"wave's frame (the closest-hit rule is order-independent: expect 0): {diff} of {W * H}")
p = hj.make_params(max_bounces=8, flags=hj.HJK_RENDER_EXACT_TIES)
frames = []
for wave in (64 << 20, 2**31 - 1):
    ctx.set_option("wave_paths", wave)
    ctx.frame_begin(W, H); ctx.render(blocks[: 135 * 64], p); frames.append(ctx.readback(normalise=False))
print("exact-tie mode, 64 spp, two wave sizes: identical frame:", np.array_equal(frames[0].view(np.uint32), frames[1].view(np.uint32)),
      "unresolved", ctx.get_info("unresolved_ties"))
