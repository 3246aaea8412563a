"""Step-by-step probe of the library's multi-GPU paths with a watchdog (prints where it is; dumps stacks if stuck).
    python tools/multi_probe.py procs   # two processes, hjk_comm_init
    python tools/multi_probe.py group   # one process, hjk_create over two devices"""
import faulthandler
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CBOX = os.path.join(ROOT, "scenes", "cbox", "cbox.obj")


def say(rank, msg):
    print(f"[{time.strftime('%H:%M:%S')}] rank {rank}: {msg}", flush=True)


def worker(rank, world, conn):
    faulthandler.dump_traceback_later(90, exit=True)
    import numpy as np
    import hijiki_b200 as hj
    if rank == 0:
        uid = hj.comm_unique_id()
        conn.send(uid)
    else:
        uid = conn.recv()
    say(rank, "have id")
    ctx = hj.Context(rank)
    ctx.comm_init(uid, rank, world)
    say(rank, "comm_init done")
    compiled = hj.Scene.from_obj(CBOX).compile()
    ctx.scene_upload(compiled)
    say(rank, f"scene uploaded, nodes {ctx.get_info('bvh_nodes')}")
    gen = hj.ImageBlockGenerator(160, 96, 64, 4)
    ctx.frame_begin(160, 96)
    ctx.render(hj.split_passes(gen.blocks(), gen.blocks_per_pass, rank, world), hj.make_params(max_bounces=8))
    say(rank, "rendered")
    ms = ctx.reduce_frame(-1)
    say(rank, f"allreduce {ms:.3f} ms")
    acc = ctx.readback(normalise=False)
    say(rank, f"readback mean {float(acc.mean()):.5f}")
    ctx.frame_begin(160, 96)
    out = ctx.readback_root(0, normalise=False)
    say(rank, f"readback_root -> {None if out is None else out.shape}")
    ctx.close()
    say(rank, "closed")


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "procs"
    if mode == "procs":
        c = mp.get_context("spawn")
        a, b = c.Pipe()
        ps = [c.Process(target=worker, args=(r, 2, a if r == 0 else b), daemon=True) for r in range(2)]
        for p in ps:
            p.start()
        for p in ps:
            p.join(timeout=150)
        print("exit codes", [p.exitcode for p in ps], flush=True)
        for p in ps:
            if p.is_alive():
                p.kill()
    else:
        faulthandler.dump_traceback_later(120, exit=True)
        import hijiki_b200 as hj
        ctx = hj.Context([0, 1])
        say("g", f"group of {ctx.get_info('n_devices')}")
        ctx.scene_upload(hj.Scene.from_obj(CBOX).compile())
        say("g", "scene uploaded")
        ctx.frame_begin(160, 96)
        st = ctx.render(hj.ImageBlockGenerator(160, 96, 64, 4).blocks(), hj.make_params(max_bounces=8))
        say("g", f"rendered {st.n_rays} rays")
        acc = ctx.readback(normalise=False)
        say("g", f"readback mean {float(acc.mean()):.5f}")
        ctx.close()
        say("g", "closed")


if __name__ == "__main__":
    main()
