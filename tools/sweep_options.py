"""Scratch perf sweep (not part of the product): kernel breakdown of cbox1080 under option settings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, json, itertools
import numpy as np
import hijiki_b200 as hj
W,H=1920,1080
scene=hj.Scene.from_obj('scenes/cbox/cbox.obj').compile()
ctx=hj.Context(0); ctx.scene_upload(scene); ctx.set_profiling(True)
blocks=hj.ImageBlockGenerator(W,H,128,32).blocks()
p=hj.make_params(max_bounces=8)
h=ctx.blocks_upload(blocks)
def run(label):
    ctx.frame_begin(W,H)
    ctx.render_resident(h,0,blocks.size,p)
    best=None
    for _ in range(3):
        ctx.frame_begin(W,H)
        st=ctx.render_resident(h,0,blocks.size,p)
        if best is None or st.ms_total<best.ms_total: best=st
    print(f"{label:34s} {best.mrays_per_s:8.0f} Mrays/s  total {best.ms_total:7.2f} ms ", {k:round(v,2) for k,v in best.kernel_ms.items()}, flush=True)
run('default')
for ft in (0,12,16,24,28,32):
    ctx.set_option('fetch_threshold',ft); run(f'fetch_threshold={ft}')
ctx.set_option('fetch_threshold',-1)
for pl in (0,4,12,16,20):
    ctx.set_option('postpone_lanes',pl); run(f'postpone_lanes={pl}')
ctx.set_option('postpone_lanes',8)
for b in (4,6,8,10):
    ctx.set_option('blocks_per_sm_traverse',b); run(f'blocks_per_sm_traverse={b}')
ctx.set_option('blocks_per_sm_traverse',8)
for b in (2,3,4,6):
    ctx.set_option('blocks_per_sm_tile',b); run(f'blocks_per_sm_tile={b}')
ctx.set_option('blocks_per_sm_tile',4)
for wp in (2<<20, 4<<20, 8<<20, 16<<20, 32<<20, 64<<20):
    ctx.set_option('wave_paths',wp); run(f'wave_paths={wp>>20}M')
