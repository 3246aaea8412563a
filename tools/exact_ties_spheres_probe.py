"""Scratch: exact-tie mode on the 512-sphere lattice: unresolved clusters and texels differing from the oracle."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import hijiki_b200 as hj
import _libs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
W, H, bs, spp, bounces = 240, 136, 64, 2, 8
compiled = hj.Scene.spheres(n).compile(use_bvh=True)
ctx = hj.Context(0); ctx.scene_upload(compiled)
blocks = hj.ImageBlockGenerator(W, H, bs, spp).blocks()
O = _libs.oracle()
acc_o = np.zeros((H, W, 4), np.float32); st = _libs.OrcStats()
op = _libs.orc_params(max_bounces=bounces, use_bvh=2, block_size=bs)
t0 = time.time()
assert O.orc_render(C.byref(compiled.view), _libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(acc_o), C.byref(st), 0) == 0
print(f"oracle {time.time()-t0:.1f}s rays {st.n_extension_rays}+{st.n_shadow_rays}")
for flags, name in ((0, "default"), (hj.HJK_RENDER_EXACT_TIES, "exact")):
    ctx.frame_begin(W, H)
    s = ctx.render(blocks, hj.make_params(max_bounces=bounces, flags=flags))
    acc = ctx.readback(normalise=False)
    diff = (acc.view(np.uint32) != acc_o.view(np.uint32)).any(axis=2)
    print(f"{name:8s} rays {s.n_extension_rays}+{s.n_shadow_rays} unresolved {ctx.get_info('unresolved_ties')} texels differing {int(diff.sum())}"
          f" nan texels gpu {int(np.isnan(acc).any(axis=2).sum())} oracle {int(np.isnan(acc_o).any(axis=2).sum())}")
