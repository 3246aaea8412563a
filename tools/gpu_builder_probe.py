"""Scratch: build time and traversal throughput of the GPU BVH builder vs the host builder."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hijiki_b200 as hj
which = sys.argv[1] if len(sys.argv) > 1 else 'terrain'
if which == 'cbox': scene = hj.Scene.from_obj('scenes/cbox/cbox.obj'); W, H, spp = 1920, 1080, 16
elif which == 'spheres': scene = hj.Scene.spheres(8); W, H, spp = 3840, 2160, 4
else: scene = hj.Scene.terrain(2237); W, H, spp = 1920, 1080, 8
compiled = scene.compile()
blocks = hj.ImageBlockGenerator(W, H, 128, spp).blocks()
p = hj.make_params(max_bounces=8)
for builder in (0, 1):
    ctx = hj.Context(0); ctx.set_profiling(True)
    ctx.set_option('bvh_builder', builder)
    t0 = time.perf_counter(); ctx.scene_upload(compiled); t1 = time.perf_counter()
    h = ctx.blocks_upload(blocks)
    best = None
    for _ in range(3):
        ctx.frame_begin(W, H); st = ctx.render_resident(h, 0, blocks.size, p)
        if best is None or st.ms_total < best.ms_total: best = st
    print(f"{which} builder={'gpu' if builder else 'host'}: upload+build {t1 - t0:.3f} s (gpu kernels {ctx.get_info('bvh_build_us') / 1e3:.1f} ms), "
          f"nodes {ctx.get_info('bvh_nodes')}, depth {ctx.get_info('bvh_depth')}, {best.mrays_per_s:.0f} Mrays/s, trace {best.kernel_ms['extend']:.2f} ms", flush=True)
    ctx.close()
