#!/usr/bin/env python
"""The reference's command line (reference src/main.rs:1426-1456) on top of hijiki_b200:

    python examples/render.py [--put-cbox-spheres] [--use-bvh] [-w 800] [-H 600] [-s 64] [-o out.exr] scene.obj

`--use-bvh` selected scene.glsl's USE_BVH walk in the reference; here every render walks the
library's own wide BVH, the flag only asks the host front-end to also emit the reference-layout
`bvh` binding.  `--present-interval` drove the preview window and is accepted for compatibility.
Prints the reference's summary line with rays over ALL bounces and device-synchronised time
(the reference counts primary rays and times asynchronous submission, src/main.rs:1487-1492).
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hijiki_b200 as hj


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--put-cbox-spheres", action="store_true")
    ap.add_argument("--use-bvh", action="store_true")
    ap.add_argument("-w", "--width", type=int, default=800)
    ap.add_argument("-H", "--height", type=int, default=600)
    ap.add_argument("--present-interval", type=int, default=128)
    ap.add_argument("-s", "--sample-count", type=int, default=64)
    ap.add_argument("-o", "--output-image", default="/tmp/output.exr")
    ap.add_argument("--max-bounces", type=int, default=1000, help="render.glsl:92")
    ap.add_argument("--exact-ties", action="store_true", help="HJK_RENDER_EXACT_TIES")
    ap.add_argument("scene")
    a = ap.parse_args()

    scene = hj.Scene.from_obj(a.scene, put_cbox_spheres=a.put_cbox_spheres)
    gen = hj.ImageBlockGenerator(a.width, a.height, 128, a.sample_count)  # block_size 128, src/main.rs:1485
    r = hj.Renderer.new(scene, gen, a.present_interval, a.use_bvh, max_bounces=a.max_bounces)
    if a.exact_ties:
        r.params.flags |= hj.HJK_RENDER_EXACT_TIES
    print("Starting rendering")
    t0 = time.perf_counter()
    st = r.render()
    dt = time.perf_counter() - t0
    print(f"Integrated {st.n_paths} paths = {st.n_rays} rays over all bounces in {dt:.3f} s "
          f"({st.n_rays / dt:.0f} rays/s; device {st.ms_total:.1f} ms, {st.n_launches} launches)")
    r.save_image(a.output_image)
    print("wrote", a.output_image)


if __name__ == "__main__":
    main()
