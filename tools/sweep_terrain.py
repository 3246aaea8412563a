"""Scratch perf sweep (not part of the product): the 10 M-triangle terrain at 1080p under trace options."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hijiki_b200 as hj
W, H = 1920, 1080
scene = hj.Scene.terrain(2237).compile()
ctx = hj.Context(0); ctx.scene_upload(scene); ctx.set_profiling(True)
blocks = hj.ImageBlockGenerator(W, H, 128, 8).blocks()
p = hj.make_params(max_bounces=8)
h = ctx.blocks_upload(blocks)
def run(label):
    ctx.frame_begin(W, H)
    ctx.render_resident(h, 0, blocks.size, p)
    best = None
    for _ in range(2):
        ctx.frame_begin(W, H)
        st = ctx.render_resident(h, 0, blocks.size, p)
        if best is None or st.ms_total < best.ms_total: best = st
    print(f"{label:34s} {best.mrays_per_s:8.0f} Mrays/s  total {best.ms_total:7.2f} ms ", {k: round(v, 2) for k, v in best.kernel_ms.items() if v}, flush=True)
run('default')
for cb in (0, 100, 140, 260, 400, 100000):
    ctx.set_option('coop_batch_cost', cb); run(f'coop_batch_cost={cb}')
ctx.set_option('coop_batch_cost', -1)
for ft in (12, 16, 24, 28, 32):
    ctx.set_option('fetch_threshold', ft); run(f'fetch_threshold={ft}')
ctx.set_option('fetch_threshold', -1)
for b in (6, 10, 12, 16):
    ctx.set_option('blocks_per_sm_traverse', b); run(f'blocks_per_sm_traverse={b}')
ctx.set_option('blocks_per_sm_traverse', 0)
ctx.set_option('coop_trace', 0); run('coop_trace=0 (per-lane k_trace)')
for pl in (0, 4, 16):
    ctx.set_option('postpone_lanes', pl); run(f'coop_trace=0 postpone_lanes={pl}')
