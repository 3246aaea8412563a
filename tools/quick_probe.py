"""Scratch: quick throughput probe for a library variant (HIJIKI_B200_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, os, time
import hijiki_b200 as hj
which=sys.argv[1] if len(sys.argv)>1 else 'cbox'
if which=='cbox': scene=hj.Scene.from_obj('scenes/cbox/cbox.obj'); W,H,spp=1920,1080,32
elif which=='spheres': scene=hj.Scene.spheres(8); W,H,spp=3840,2160,4
else: scene=hj.Scene.terrain(2237); W,H,spp=1920,1080,8
ctx=hj.Context(0); ctx.scene_upload(scene.compile()); ctx.set_profiling(True)
for kv in sys.argv[2:]:
    k,v=kv.split('='); ctx.set_option(k,int(v))
blocks=hj.ImageBlockGenerator(W,H,128,spp).blocks()
p=hj.make_params(max_bounces=8)
h=ctx.blocks_upload(blocks)
best=None
for _ in range(4):
    ctx.frame_begin(W,H)
    st=ctx.render_resident(h,0,blocks.size,p)
    if best is None or st.ms_total<best.ms_total: best=st
print(f"{os.environ.get('HIJIKI_B200_LIB','default'):45s} {which:8s} {' '.join(sys.argv[2:]):20s} {best.mrays_per_s:8.0f} Mrays/s total {best.ms_total:7.2f} ms", {k:round(v,2) for k,v in best.kernel_ms.items() if v>0}, flush=True)
