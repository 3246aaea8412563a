#!/usr/bin/env python
"""bench.py — Mrays/s (all bounces) of the path-tracing hot path on BASELINE.json's config 2
(scenes/cbox, 1920x1080, 1024 spp, max 8 bounces), plus the reconstruction ("denoiser") HBM GB/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One STEP = `--spp-per-step` consecutive sample passes of the 1024-spp job on every rank (a
slice of the job's own block list: 135 ImageBlocks per pass), integrated, reconstructed into the
accumulator and — for N > 1 — all-reduced over NCCL.  One RAY = one intersectScene call
(reference shader/render.glsl:94,122): extension + shadow rays.

  value      device-resident throughput: the step's block list already lives in HBM; timed with
             CUDA events on the stream the kernels run on, max over ranks.
  e2e        the same step through the public call (`hjk_render` with a HOST block list in pinned
             memory + `hjk_readback` of the accumulator to pinned host memory), copies in the timed
             region.
  roofline   the dominant kernel (k_extend, closest-hit BVH traversal): algorithmic bytes it must
             move per ray / its average launch time, against the measured HBM copy peak.
  cpu_baseline / --impl reference
             the CPU restatement of the reference GLSL (oracle/, threaded-BVH2 mode = the
             reference's --use-bvh), on all host cores, on a bounded sample of the same workload.
             The reference's own wgpu/lavapipe path cannot run in this image.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SPP, MAX_BOUNCES, BLOCK = 1920, 1080, 1024, 8, 128
CBOX = os.path.join(ROOT, "scenes", "cbox", "cbox.obj")
WORKLOAD = "scenes/cbox 1920x1080, 1024 spp, max 8 bounces (BASELINE.json configs[1])"
# algorithmic bytes: what k_extend itself must move per ray, and the whole
# pipeline's per-ray queue traffic of SURVEY.md §8(d)
EXTEND_BYTES_PER_RAY = 4 + 32 + 16
PIPE_BYTES_EXT, PIPE_BYTES_SHADOW = 192, 96
RECON_BYTES_PER_PX = 64  # 2 x 16 B layers + accumulator read + write (the all-zero albedo layer is elided)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def job_blocks():
    import hijiki_b200 as hj
    gen = hj.ImageBlockGenerator(WIDTH, HEIGHT, BLOCK, SPP)
    return gen, gen.blocks()


def step_slice(blocks, bpp, step, rank, world, spp_per_step):
    """Blocks of the passes rank `rank` renders in step `step`: pass p -> rank p mod world."""
    n_pass = len(blocks) // bpp
    idx = []
    for i in range(spp_per_step):
        p = ((step * spp_per_step + i) * world + rank) % n_pass
        idx.append(np.arange(p * bpp, (p + 1) * bpp))
    return np.ascontiguousarray(blocks[np.concatenate(idx)])


# ------------------------------------------------------------------------------ reference arm
def run_reference(args, rank):
    """The reference's algorithm on the host CPU (oracle port, all cores)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _libs
    import hijiki_b200 as hj

    O = _libs.oracle()
    cores = O.orc_hardware_threads()
    compiled = hj.Scene.from_obj(CBOX).compile(use_bvh=True)
    gen, blocks = job_blocks()
    bpp = gen.blocks_per_pass
    op = _libs.orc_params(max_bounces=MAX_BOUNCES, use_bvh=1, block_size=BLOCK)
    acc = np.zeros((HEIGHT, WIDTH, 4), np.float32)

    def run(sample):
        st = _libs.OrcStats()
        t0 = time.perf_counter()
        rc = O.orc_render(C.byref(compiled.view), _libs.ptr(sample), sample.size, C.byref(op), _libs.ptr(acc),
                          C.byref(st), cores)
        dt = time.perf_counter() - t0
        assert rc == 0
        return st.n_extension_rays + st.n_shadow_rays, dt

    # calibrate on two blocks, then size a step so the whole run stays within ~100 s
    rays, dt = run(blocks[60:62])
    per_block = max(dt / 2, 1e-3)
    budget = 100.0 / max(args.steps + args.warmup, 1)
    nb = int(min(32 * bpp, max(2, budget / per_block)))
    n_total = len(blocks)
    for w in range(args.warmup):
        run(blocks[(w * nb) % (n_total - nb):][:nb])
    tot_rays, tot_dt = 0, 0.0
    for k in range(args.steps):
        first = ((args.warmup + k) * nb) % (n_total - nb)
        r, dt = run(blocks[first:first + nb])
        tot_rays += r
        tot_dt += dt
    value = tot_rays / tot_dt / 1e6
    sample = f"{nb} ImageBlocks ({nb / bpp:.1f} sample passes of 1920x1080) per step"
    line = {
        "impl": "reference", "metric": "Mrays/s (all bounces)", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement of reference GLSL (threaded-BVH2 mode), "
                   f"{cores} cores — the reference's wgpu/lavapipe path cannot run in this image"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(compiled, blocks, bpp):
    """Bounded CPU sample for the main line (rank 0, N = 1): ~10-20 s of oracle work."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _libs
    O = _libs.oracle()
    cores = O.orc_hardware_threads()
    op = _libs.orc_params(max_bounces=MAX_BOUNCES, use_bvh=1, block_size=BLOCK)
    acc = np.zeros((HEIGHT, WIDTH, 4), np.float32)
    st = _libs.OrcStats()
    t0 = time.perf_counter()
    O.orc_render(C.byref(compiled.view), _libs.ptr(blocks[60:62]), 2, C.byref(op), _libs.ptr(acc), C.byref(st), cores)
    per_block = max((time.perf_counter() - t0) / 2, 1e-4)
    nb = int(min(len(blocks), max(2, 15.0 / per_block)))  # ~15 s of CPU work
    st = _libs.OrcStats()
    t0 = time.perf_counter()
    O.orc_render(C.byref(compiled.view), _libs.ptr(blocks[:nb]), nb, C.byref(op), _libs.ptr(acc), C.byref(st), cores)
    dt = time.perf_counter() - t0
    rays = st.n_extension_rays + st.n_shadow_rays
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"first {nb} ImageBlocks ({nb / bpp:.1f} sample passes of 1920x1080) of the job, "
                      f"{rays / 1e6:.1f} Mrays in {dt:.1f} s (oracle, threaded-BVH2 mode = reference --use-bvh)"}


# ------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-denoiser", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import hijiki_b200 as hj

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    gen, blocks = job_blocks()
    bpp = gen.blocks_per_pass
    compiled = hj.Scene.from_obj(CBOX).compile(use_bvh=True)
    ctx = hj.Context(local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    ctx.scene_upload(compiled)
    ctx.frame_begin(WIDTH, HEIGHT)
    ctx.set_profiling(True)
    params = hj.make_params(max_bounces=MAX_BOUNCES)
    n_steps_total = args.warmup + args.steps

    # accumulator as a torch tensor (no copy) so torch.distributed can reduce it in place
    acc_ptr, acc_n = ctx.accumulator_device_ptr()

    class _Alias:
        __cuda_array_interface__ = {"shape": (acc_n,), "typestr": "<f4", "data": (acc_ptr, False), "version": 2}

    acc_t = torch.as_tensor(_Alias(), device=torch.device("cuda", local_rank))

    def allreduce():
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(acc_t)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm: block lists uploaded before the timed region
    slices = [step_slice(blocks, bpp, s, rank, world, args.spp_per_step) for s in range(n_steps_total)]
    handles = [ctx.blocks_upload(s) for s in slices]
    for s in range(args.warmup):
        ctx.render_resident(handles[s], 0, slices[s].size, params)
        allreduce()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays = ext_rays = sh_rays = launches = 0
    kernel_ms = {}
    ev0.record(stream)
    for s in range(args.warmup, n_steps_total):
        st = ctx.render_resident(handles[s], 0, slices[s].size, params)
        allreduce()
        rays += st.n_rays
        ext_rays += st.n_extension_rays
        sh_rays += st.n_shadow_rays
        launches += st.n_launches + (1 if world > 1 else 0)
        for k, v in st.kernel_ms.items():
            kernel_ms[k] = kernel_ms.get(k, 0.0) + v
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    for h in handles:
        ctx.blocks_free(h)

    # ---------------- end-to-end arm: host block list in, accumulator out, every step
    pinned_blocks = [torch.from_numpy(s.view(np.uint8).copy()).pin_memory() for s in slices]
    out_host = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory()
    for s in range(args.warmup):
        ctx.render((pinned_blocks[s].data_ptr(), slices[s].size), params)
        allreduce()
        ctx.readback_ptr(out_host.data_ptr(), WIDTH * 16, normalise=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_rays = 0
    e0.record(stream)
    for s in range(args.warmup, n_steps_total):
        st = ctx.render((pinned_blocks[s].data_ptr(), slices[s].size), params)
        allreduce()
        ctx.readback_ptr(out_host.data_ptr(), WIDTH * 16, normalise=True)
        e2e_rays += st.n_rays
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    result_mean = float(out_host[..., :3].mean())

    # ---------------- reconstruction ("denoiser") bandwidth on 3840x2160 feature buffers (config 5)
    denoiser = None
    if not args.no_denoiser:
        dw, dh = 3840, 2160
        rng = np.random.default_rng(5)
        rad = np.exp(rng.standard_normal((dh, dw, 4), dtype=np.float32))
        rad[..., 3] = 1.0
        nrm = rng.standard_normal((dh, dw, 4), dtype=np.float32)
        nrm[..., :3] /= np.linalg.norm(nrm[..., :3], axis=2, keepdims=True)
        dblocks = hj.ImageBlockGenerator(dw, dh, BLOCK, 1).blocks()
        dctx = hj.Context(local_rank)
        dctx.frame_begin(dw, dh)
        dctx.denoise_upload(rad, nrm, dblocks)
        dctx.denoise_resident(params, 5)
        reps = 50
        dms = dctx.denoise_resident(params, reps)
        gbs = RECON_BYTES_PER_PX * dw * dh * reps / (dms * 1e-3) / 1e9
        denoiser = {"workload": "3840x2160 synthetic feature buffers, 510 ImageBlocks/pass, R=2", "ms_per_pass": dms / reps,
                    "bytes_per_px": RECON_BYTES_PER_PX, "achieved": gbs, "unit": "GB/s"}
        dctx.close()

    # ---------------- gather over ranks (max time, summed rays)
    peak, peak_src = measured_peaks()
    if world > 1:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        c = torch.tensor([rays, e2e_rays, ext_rays, sh_rays, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(c)
        rays, e2e_rays, ext_rays, sh_rays, launches = (int(v) for v in c.tolist())
    if rank == 0:
        value = rays / (ms * 1e-3) / 1e6
        e2e = e2e_rays / (e2e_ms * 1e-3) / 1e6
        n_ext_launches = args.steps * MAX_BOUNCES * max(1, -(-args.spp_per_step // max(1, round((4 << 20) / (WIDTH * HEIGHT)))))
        ext_ms = kernel_ms.get("extend", 0.0)
        ext_local = ext_rays // world
        achieved = EXTEND_BYTES_PER_RAY * ext_local / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else None
        line = {
            "metric": "Mrays/s (all bounces)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "spp_per_step_per_gpu": args.spp_per_step,
                       "blocks_per_step_per_gpu": args.spp_per_step * bpp, "parallelism": f"sample-pass dp{world}",
                       "l2": "per-step working set (path state + queues of a 4M-path wave, ~0.7 GB) exceeds the 126 MB L2",
                       "image_mean": result_mean},
            "rays": {"extension": ext_rays, "shadow": sh_rays, "per_path": rays / max(1, args.steps * args.spp_per_step * WIDTH * HEIGHT * world)},
            "e2e": {"value": e2e, "unit": "Mrays/s", "h2d_bytes_per_step": int(slices[0].nbytes),
                    "d2h_bytes_per_step": WIDTH * HEIGHT * 16, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "kernel_ms_per_step": {k: v / args.steps for k, v in kernel_ms.items()},
            "roofline": {"kernel": "k_extend", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                         "bytes_per_ray": EXTEND_BYTES_PER_RAY, "launches": n_ext_launches,
                         "avg_launch_ms": ext_ms / n_ext_launches if n_ext_launches else None,
                         "share_of_step": ext_ms / ms if ms else None,
                         "note": "cbox (0.35 MB of BVH) is cache-resident: traversal is issue/latency-bound, not HBM-bound"},
            "pipeline_bytes": {"achieved": (PIPE_BYTES_EXT * ext_rays + PIPE_BYTES_SHADOW * sh_rays) / world / (ms * 1e-3) / 1e9,
                               "unit": "GB/s", "frac": (PIPE_BYTES_EXT * ext_rays + PIPE_BYTES_SHADOW * sh_rays) / world / (ms * 1e-3) / 1e9 / peak},
            "clocks": clocks,
        }
        if denoiser:
            denoiser["peak"] = peak
            denoiser["frac"] = denoiser["achieved"] / peak
            line["denoiser"] = denoiser
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample(compiled, blocks, bpp)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
