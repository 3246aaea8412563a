#!/usr/bin/env python
"""bench.py — Mrays/s (all bounces) of the path-tracing hot path, plus the reconstruction
("denoiser") HBM GB/s.  Default workload: BASELINE.json configs[1] (scenes/cbox, 1920x1080,
1024 spp, max 8 bounces).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One STEP = one frame of `--spp-per-step` consecutive sample passes of the job on every rank (a
slice of the job's own ImageBlock list): accumulator zeroed, passes integrated and reconstructed
into it and — for N > 1 — all-reduced over NCCL.  One RAY = one intersectScene call
(reference shader/render.glsl:94,122): extension + shadow rays.

  value      device-resident throughput: the step's block list already lives in HBM; timed with
             CUDA events on the stream the kernels run on, max over ranks.
  e2e        the same step through the public call (`hjk_render` with a HOST block list in pinned
             memory + `hjk_readback` of the normalised frame to pinned host memory), copies in the
             timed region.
  roofline   the dominant kernel (k_trace_coop, BVH traversal of extension + shadow rays): algorithmic
             bytes it must move per ray / its average launch time, against the measured HBM copy peak.
  cpu_baseline / --impl reference
             the CPU restatement of the reference GLSL (oracle/, threaded-BVH2 mode = the
             reference's --use-bvh), on all host cores, on a bounded sample of the same workload.
             The reference's own wgpu/lavapipe path cannot run in this image.

Other workloads (`--workload`): cbox_default (configs[0]), terrain (configs[2], 10 M triangles),
spheres (configs[3], 512 dielectric/mirror spheres at 3840x2160; spheres64 = the same with max 64 bounces); configs[4] (reconstruction on
3840x2160 feature buffers) is the `denoiser` object of every line.
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

BLOCK = 128
CBOX = os.path.join(ROOT, "scenes", "cbox", "cbox.obj")
# name -> (label, scene kind, width, height, spp, max_bounces, default spp per step)
WORKLOADS = {
    "cbox1080": ("scenes/cbox 1920x1080, 1024 spp, max 8 bounces (BASELINE.json configs[1])", "cbox", 1920, 1080, 1024, 8, 32),
    "cbox_default": ("scenes/cbox 800x600, 64 spp, max 1000 bounces, reference defaults (BASELINE.json configs[0])", "cbox", 800, 600, 64, 1000, 64),
    "terrain": ("synthetic 10,008,370-triangle checkerboard terrain + emissive quads, 1920x1080, 64 spp, max 8 bounces (BASELINE.json configs[2])", "terrain", 1920, 1080, 64, 8, 8),
    "spheres": ("synthetic 512-sphere dielectric/mirror lattice, 3840x2160, 4096 spp, max 8 bounces (BASELINE.json configs[3])", "spheres", 3840, 2160, 4096, 8, 4),
    "spheres64": ("synthetic 512-sphere dielectric/mirror lattice, 3840x2160, 4096 spp, max 64 bounces (SURVEY 8d: the variant that exposes path-length divergence)", "spheres", 3840, 2160, 4096, 64, 4),
}
# algorithmic bytes: what k_trace itself must move per ray (extension: queue entry +
# ray in, hit record out; shadow: ray + payload in), and the whole pipeline's per-ray queue traffic
# of SURVEY.md §8(d)
EXTEND_BYTES_PER_RAY = 4 + 32 + 16
SHADOW_BYTES_PER_RAY = 32 + 16
PIPE_BYTES_EXT, PIPE_BYTES_SHADOW = 192, 96
RECON_BYTES_PER_PX = 64  # 2 x 16 B layers + accumulator read + write (the all-zero albedo layer is elided)
# dram__bytes_read.sum + dram__bytes_write.sum of k_trace_coop per traced ray, from the ncu --set full capture of
# one whole wave in profiles/r01f_ncu_full_wave_selected.csv (14.59 GB over its nine k_trace_coop launches,
# 216.3 M rays; 17.58 GB = 81.3 B/ray before the path state became queue-ordered)
TRACE_DRAM_BYTES_PER_RAY_NCU = 67.5


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
                                          "-lms", "50", "-i", str(self.device)], stdout=subprocess.PIPE,
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
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_scene(kind: str):
    import hijiki_b200 as hj
    if kind == "cbox":
        return hj.Scene.from_obj(CBOX)
    if kind == "terrain":
        return hj.Scene.terrain(2237)
    if kind == "spheres":
        return hj.Scene.spheres(8)
    raise KeyError(kind)


def step_slice(blocks, bpp, step, rank, world, spp_per_step):
    """Blocks of the passes rank `rank` renders in step `step`: pass p -> rank p mod world."""
    n_pass = len(blocks) // bpp
    idx = []
    for i in range(spp_per_step):
        p = ((step * spp_per_step + i) * world + rank) % n_pass
        idx.append(np.arange(p * bpp, (p + 1) * bpp))
    return np.ascontiguousarray(blocks[np.concatenate(idx)])


# ------------------------------------------------------------------------------ CPU arm
def _oracle_setup(wl):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _libs
    import hijiki_b200 as hj
    label, kind, width, height, spp, max_bounces, _ = wl
    O = _libs.oracle()
    cores = O.orc_hardware_threads()
    compiled = make_scene(kind).compile(use_bvh=True)
    gen = hj.ImageBlockGenerator(width, height, BLOCK, spp)
    # bounded block list: never materialise millions of blocks for the CPU arm
    gen.num_samples = min(spp, 64)
    blocks = gen.blocks()
    op = _libs.orc_params(max_bounces=max_bounces, use_bvh=1, block_size=BLOCK)
    acc = np.zeros((height, width, 4), np.float32)

    def run(sample):
        st = _libs.OrcStats()
        t0 = time.perf_counter()
        rc = O.orc_render(C.byref(compiled.view), _libs.ptr(sample), sample.size, C.byref(op), _libs.ptr(acc),
                          C.byref(st), cores)
        dt = time.perf_counter() - t0
        assert rc == 0
        return st.n_extension_rays + st.n_shadow_rays, dt

    return run, blocks, gen.blocks_per_pass, cores


def run_reference(args, rank, wl):
    """The reference's algorithm on the host CPU (oracle port, all cores)."""
    if rank != 0:
        return
    label, kind, width, height = wl[:4]
    base = {"impl": "reference", "metric": "Mrays/s (all bounces)", "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if kind == "terrain":
        base["unavailable"] = ("the reference's flattened BVH uses exit index 1000000 as its end sentinel "
                               "(src/main.rs:231) and cannot represent a 10 M-triangle scene")
        emit(base)
        return
    run, blocks, bpp, cores = _oracle_setup(wl)
    mid = bpp // 2
    _, dt = run(blocks[mid:mid + 2])  # calibrate on two blocks from the middle of the frame
    per_block = max(dt / 2, 1e-4)
    budget = args.cpu_seconds / max(args.steps + args.warmup, 1)
    nb = int(min(len(blocks) // 2, max(2, budget / per_block)))
    for w in range(args.warmup):
        run(blocks[(w * nb) % (len(blocks) - nb):][:nb])
    tot_rays, tot_dt = 0, 0.0
    for k in range(args.steps):
        first = ((args.warmup + k) * nb) % (len(blocks) - nb)
        r, dt = run(blocks[first:first + nb])
        tot_rays += r
        tot_dt += dt
    value = tot_rays / tot_dt / 1e6
    sample = f"{nb} ImageBlocks ({nb / bpp:.2f} sample passes of {width}x{height}) per step"
    base.update({
        "value": value, "ms_per_step": tot_dt / args.steps * 1e3,
        "config": {"workload": label, "note": "CPU restatement of reference GLSL (threaded-BVH2 mode = reference "
                   f"--use-bvh), {cores} host threads — the reference's wgpu/lavapipe path cannot run in this image"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    emit(base)


def cpu_baseline_sample(wl):
    """Bounded CPU sample for the main line (rank 0, N = 1): ~20 s of oracle work."""
    label, kind, width, height = wl[:4]
    if kind == "terrain":
        return {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "port",
                "sample": "unavailable: the reference BVH layout cannot represent 10 M triangles (src/main.rs:231)"}
    run, blocks, bpp, cores = _oracle_setup(wl)
    mid = bpp // 2
    _, dt = run(blocks[mid:mid + 2])
    per_block = max(dt / 2, 1e-4)
    nb = int(min(len(blocks), max(2, 20.0 / per_block)))
    rays, dt = run(blocks[:nb])
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"first {nb} ImageBlocks ({nb / bpp:.1f} sample passes of {width}x{height}) of the job, "
                      f"{rays / 1e6:.1f} Mrays in {dt:.1f} s (oracle, threaded-BVH2 mode = reference --use-bvh)"}


# ------------------------------------------------------------------------------ our arm
_RESULT_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    the first communicator), so everything but the result line is sent to stderr: fd 1 is pointed at fd 2 and
    the original stdout kept for emit()."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cbox1080", choices=sorted(WORKLOADS))
    ap.add_argument("--spp-per-step", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-denoiser", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=100.0, help="wall budget of the --impl reference run")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    label, kind, width, height, spp, max_bounces, default_sps = wl
    sps = args.spp_per_step or default_sps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, wl)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import hijiki_b200 as hj

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_steps_total = args.warmup + args.steps
    gen = hj.ImageBlockGenerator(width, height, BLOCK, spp)
    bpp = gen.blocks_per_pass
    # only the passes the run touches are generated (the list is a prefix of the job's list)
    gen.num_samples = min(spp, n_steps_total * sps * world)
    blocks = gen.blocks()
    t0 = time.perf_counter()
    compiled = make_scene(kind).compile(use_bvh=False)
    t_scene = time.perf_counter() - t0
    ctx = hj.Context(local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    ctx.scene_upload(compiled)
    t_upload = time.perf_counter() - t0
    ctx.frame_begin(width, height)
    ctx.set_profiling(True)
    params = hj.make_params(max_bounces=max_bounces)

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
    slices = [step_slice(blocks, bpp, s, rank, world, sps) for s in range(n_steps_total)]
    handles = [ctx.blocks_upload(s) for s in slices]
    for s in range(args.warmup):
        ctx.frame_begin(width, height)
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
        ctx.frame_begin(width, height)
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

    # exact-tie mode (HJK_RENDER_EXACT_TIES), two steps: what bit-identity with the reference on every ray costs
    exact_params = hj.make_params(max_bounces=max_bounces, flags=hj.HJK_RENDER_EXACT_TIES)
    exact_rays, exact_ms = 0, 0.0
    for s in range(args.warmup, min(n_steps_total, args.warmup + 2)):
        ctx.frame_begin(width, height)
        st = ctx.render_resident(handles[s], 0, slices[s].size, exact_params)
        exact_rays += st.n_rays
        exact_ms += st.ms_total
    exact_unresolved = ctx.get_info("unresolved_ties")
    for h in handles:
        ctx.blocks_free(h)

    # ---------------- end-to-end arm: host block list in, normalised frame out, every step
    e2e_ms, e2e_rays, result_mean = None, 0, None
    if not args.no_e2e:
        pinned_blocks = [torch.from_numpy(s.view(np.uint8).copy()).pin_memory() for s in slices]
        out_host = torch.empty((height, width, 4), dtype=torch.float32).pin_memory()

        def e2e_step(s):
            ctx.frame_begin(width, height)
            st = ctx.render((pinned_blocks[s].data_ptr(), slices[s].size), params)
            allreduce()
            ctx.readback_ptr(out_host.data_ptr(), width * 16, normalise=True)
            return st.n_rays

        for s in range(args.warmup):
            e2e_step(s)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in range(args.warmup, n_steps_total):
            e2e_rays += e2e_step(s)
        e1.record(stream)
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        result_mean = float(out_host[..., :3].mean())

    # the path's one collective, timed alone: all-reduce(sum) of the W x H float4 accumulator
    allreduce_info = None
    if world > 1:
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        allreduce()
        a0.record(stream)
        for _ in range(reps):
            allreduce()
        a1.record(stream)
        barrier()
        ar_ms = a0.elapsed_time(a1) / reps
        nbytes = width * height * 16
        allreduce_info = {"bytes": nbytes, "ms": ar_ms,
                          "bus_gbs": 2 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9,
                          "note": "torch.distributed NCCL on the render stream; bus bandwidth = 2(N-1)/N * S / t "
                                  "(nominal NVLink 5: 900 GB/s per direction)"}

    info = {k: ctx.get_info(k) for k in ("bvh_nodes", "bvh_prims", "bvh_depth", "bvh_bytes", "wave_paths",
                                          "blocks_per_sm_traverse", "blocks_per_sm_tile", "n_sms")}
    ctx.close()

    # ---------------- reconstruction ("denoiser") bandwidth on 3840x2160 feature buffers (configs[4])
    denoiser = None
    if not args.no_denoiser:
        dw, dh = 3840, 2160
        rng = np.random.default_rng(5 + rank)
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
        denoiser = {"workload": "3840x2160 synthetic feature buffers (random unit normals), 510 ImageBlocks/pass, R=2 "
                                "(BASELINE.json configs[4]), per GPU",
                    "ms_per_pass": dms / reps, "bytes_per_px": RECON_BYTES_PER_PX, "achieved": gbs, "unit": "GB/s"}
        dctx.close()

    # ---------------- gather over ranks (max time, summed rays)
    peak, peak_src = measured_peaks()
    if world > 1:
        t = torch.tensor([ms, e2e_ms or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), (float(t[1]) if e2e_ms is not None else None)
        c = torch.tensor([rays, e2e_rays, ext_rays, sh_rays, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(c)
        rays, e2e_rays, ext_rays, sh_rays, launches = (int(v) for v in c.tolist())
    if rank == 0:
        value = rays / (ms * 1e-3) / 1e6
        wave_passes = max(1, min(sps, (info["wave_paths"] + width * height // 2) // (width * height)))
        n_ext_launches = args.steps * (max_bounces + 1) * -(-sps // wave_passes)
        ext_ms = kernel_ms.get("extend", 0.0)  # slot HJK_K_EXTEND times k_trace (extension + shadow rays)
        trace_bytes = (EXTEND_BYTES_PER_RAY * ext_rays + SHADOW_BYTES_PER_RAY * sh_rays) / world
        achieved = trace_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else None
        pipe = (PIPE_BYTES_EXT * ext_rays + PIPE_BYTES_SHADOW * sh_rays) / world / (ms * 1e-3) / 1e9
        line = {
            "metric": "Mrays/s (all bounces)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": label, "spp_per_step_per_gpu": sps, "blocks_per_step_per_gpu": sps * bpp,
                       "parallelism": f"sample-pass dp{world}",
                       "l2": "inputs larger than L2: the path state + queues of one wave (~200 B per path, GBs per wave) exceed the 126 MB L2",
                       "image_mean": result_mean, "bvh": info,
                       "scene_build_s": round(t_scene, 2), "bvh_build_upload_s": round(t_upload, 2)},
            "rays": {"extension": ext_rays, "shadow": sh_rays,
                     "per_path": rays / max(1, args.steps * sps * width * height * world)},
            "gpu_launches": launches,
            "kernel_ms_per_step": {k: v / args.steps for k, v in kernel_ms.items()},
            "roofline": {"kernel": "k_trace_coop", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": (TRACE_DRAM_BYTES_PER_RAY_NCU * rays / world / n_ext_launches
                                     if kind == "cbox" and n_ext_launches else None),
                         "algorithmic_bytes_per_launch": trace_bytes / n_ext_launches if n_ext_launches else None,
                         "traffic_source": "ncu DRAM bytes per traced ray (profiles/r01f, cbox) x rays per launch",
                         "peak_source": peak_src,
                         # what actually bounds traversal on a cache-resident scene (SURVEY 8d caveat): cycles with an
                         # instruction issued / lanes per instruction, from the same ncu capture (not measured live)
                         "issue_active_pct_ncu": 74.2 if kind == "cbox" else None,
                         "lanes_per_instruction_ncu": 21.2 if kind == "cbox" else None,
                         "bytes_per_ray": {"extension": EXTEND_BYTES_PER_RAY, "shadow": SHADOW_BYTES_PER_RAY},
                         "launches": n_ext_launches,
                         "avg_launch_ms": ext_ms / n_ext_launches if n_ext_launches else None,
                         "share_of_step": ext_ms / ms if ms else None,
                         "note": "traversal of a cache-resident BVH is issue/latency-bound, not HBM-bound "
                                 "(SURVEY.md §8d); see profiles/ for SM issue utilisation"},
            "pipeline_bytes": {"achieved": pipe, "unit": "GB/s", "frac": pipe / peak},
            "clocks": clocks,
            "exact_ties": {"value": exact_rays / (exact_ms * 1e-3) / 1e6 if exact_ms else None, "unit": "Mrays/s",
                           "unresolved_clusters_last_step": exact_unresolved,
                           "note": "this rank, 2 steps, HJK_RENDER_EXACT_TIES (the mode whose frames the parity tests "
                                   "hold bit-identical to the oracle); unresolved = tie clusters longer than the "
                                   "8*M_EPS window or 12 candidates, counted, each can move one sample"},
        }
        if e2e_ms is not None:
            line["e2e"] = {"value": e2e_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s",
                           "h2d_bytes_per_step": int(slices[0].nbytes), "d2h_bytes_per_step": width * height * 16,
                           "ms_per_step": e2e_ms / args.steps}
        if allreduce_info:
            line["allreduce"] = allreduce_info
        if denoiser:
            denoiser["peak"] = peak
            denoiser["frac"] = denoiser["achieved"] / peak
            line["denoiser"] = denoiser
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample(wl)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
