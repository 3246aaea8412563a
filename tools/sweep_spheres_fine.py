import os, sys
sys.path.insert(0, os.getcwd())
import hijiki_b200 as hj
W, H = 3840, 2160
scene = hj.Scene.spheres(8).compile()
ctx = hj.Context(0); ctx.scene_upload(scene); ctx.set_profiling(True)
blocks = hj.ImageBlockGenerator(W, H, 128, 4).blocks()
h = ctx.blocks_upload(blocks)
for mb in (64, 8):
    p = hj.make_params(max_bounces=mb)
    for cb in (220, 260, 300):
        for ft in (22, 24, 26):
            ctx.set_option('coop_batch_cost', cb); ctx.set_option('fetch_threshold', ft)
            best = None
            for _ in range(3):
                ctx.frame_begin(W, H)
                st = ctx.render_resident(h, 0, blocks.size, p)
                if best is None or st.ms_total < best.ms_total: best = st
            print(f"bounces {mb} cost {cb} fetch {ft}: {best.mrays_per_s:8.0f} Mrays/s extend {best.kernel_ms['extend']:.2f}", flush=True)
