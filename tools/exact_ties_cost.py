"""Scratch: cost of the exact-tie mode on cbox1080."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hijiki_b200 as hj
W,H,spp=1920,1080,32
ctx=hj.Context(0); ctx.scene_upload(hj.Scene.from_obj('scenes/cbox/cbox.obj').compile()); ctx.set_profiling(True)
blocks=hj.ImageBlockGenerator(W,H,128,spp).blocks()
h=ctx.blocks_upload(blocks)
for flags,name in ((0,'default'),(hj.HJK_RENDER_EXACT_TIES,'exact ties')):
    p=hj.make_params(max_bounces=8, flags=flags)
    best=None
    for _ in range(4):
        ctx.frame_begin(W,H); st=ctx.render_resident(h,0,blocks.size,p)
        if best is None or st.ms_total<best.ms_total: best=st
    print(f"{name:12s} {best.mrays_per_s:8.0f} Mrays/s total {best.ms_total:7.2f} ms trace {best.kernel_ms['extend']:.2f} unresolved {ctx.get_info('unresolved_ties')}")
