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
             memory + `hjk_readback_root` of the normalised frame to pinned host memory), copies in the
             timed region.  N > 1: the library's own communicator (hjk_comm_init, bootstrapped over the
             torchrun rendezvous) sums the frames with one ncclReduce to rank 0, which alone reads back.
  roofline   the dominant kernel (k_trace_coop, BVH traversal of extension + shadow rays): algorithmic
             bytes it must move per ray / its average launch time, against the measured HBM copy peak.
  cpu_baseline / --impl reference
             the CPU restatement of the reference GLSL (oracle/, threaded-BVH2 mode = the
             reference's --use-bvh), on all host cores, on a bounded sample of the same workload.
             The reference's own wgpu/lavapipe path cannot run in this image.

The default line also carries, as objects measured the same way (fewer steps), the other configs of
BASELINE.json: `terrain` (configs[2], 10 M triangles), `spheres` (configs[3], 512 dielectric/mirror spheres at
3840x2160; `spheres64` = the same with max 64 bounces), `denoiser` (configs[4], reconstruction on 3840x2160
feature buffers), and `strong`: a fixed 256-spp cbox 1080p job split over the N GPUs, timed from scene upload
(BVH built once, broadcast) to the frame on rank 0's host.  `--workload NAME` makes one of them the main line.
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
SCENE_FETCH_BYTES_PER_RAY = 784  # uncached lower bound for scenes exceeding L2 (terrain): 8 nodes x 80 + 3 x 48
NCU_METRICS = os.path.join(ROOT, "profiles", "ncu_trace_metrics.json")
KERNEL_SOURCES = ("kernels.cuh", "traverse.cuh", "shade.cuh", "recon.cuh", "hjk_math.cuh", "scene_dev.cuh")


def kernel_source_sha() -> str:
    """Digest of the device sources the ncu-derived figures belong to."""
    import hashlib
    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "hijiki_b200", "csrc", "device", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_metrics(kind: str):
    """Figures that only a profiler can give (DRAM bytes per traced ray, issue-active cycles, lanes per instruction of
    the trace kernel), read from the committed summary of the ncu capture — written by tools/ncu_trace_metrics.py from
    the capture's raw page — never hard-coded here.  `stale` says the device sources changed since the capture."""
    try:
        with open(NCU_METRICS) as f:
            m = json.load(f)
    except Exception:
        return None
    e = m.get("workloads", {}).get(kind)
    if not e:
        return None
    e = dict(e)
    e["source"] = m.get("source")
    e["kernel_source_sha"] = m.get("kernel_source_sha")
    e["stale"] = m.get("kernel_source_sha") != kernel_source_sha()
    return e


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
# Self-contained: scene arrays, the reference-layout BVH2 and the block list come from the TESTS-side build of the
# host front-end (tests/native/libhjk_hosttest.so), the path tracer is oracle/liboracle.so.  The product's
# libhijiki_b200.so is never loaded by this arm.
def _host_scene(_libs, kind):
    H = _libs.hosttest()
    if kind == "cbox":
        return _libs.HostScene.from_obj(H, CBOX, put_spheres=False, with_bvh2=True)
    if kind == "terrain":
        return _libs.HostScene.terrain(H, 2237, with_bvh2=False)
    if kind == "spheres":
        return _libs.HostScene.spheres(H, 8, with_bvh2=True)
    raise KeyError(kind)


def _oracle_setup(wl):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _libs
    label, kind, width, height, spp, max_bounces, _ = wl
    O = _libs.oracle()
    cores = O.orc_hardware_threads()
    scene = _host_scene(_libs, kind)
    # bounded block list: never materialise millions of blocks for the CPU arm
    blocks = _libs.generate_blocks(_libs.hosttest(), width, height, min(spp, 64), BLOCK)
    bpp = ((width + BLOCK - 1) // BLOCK) * ((height + BLOCK - 1) // BLOCK)
    # threaded BVH2 = what the reference's --use-bvh walks; the 10 M-triangle terrain cannot be represented in that
    # layout (exit sentinel 1000000, src/main.rs:231): it runs the oracle's culled linear scan instead
    mode = 3 if kind == "terrain" else 1
    op = _libs.orc_params(max_bounces=max_bounces, use_bvh=mode, block_size=BLOCK)
    acc = np.zeros((height, width, 4), np.float32)

    def run(sample):
        st = _libs.OrcStats()
        t0 = time.perf_counter()
        rc = O.orc_render(C.byref(scene.view), _libs.ptr(sample), sample.size, C.byref(op), _libs.ptr(acc),
                          C.byref(st), cores)
        dt = time.perf_counter() - t0
        assert rc == 0
        return st.n_extension_rays + st.n_shadow_rays, dt

    how = ("oracle, culled linear scan (mode 3): the reference's BVH layout cannot hold 10 M triangles" if mode == 3
           else "oracle, threaded-BVH2 mode = reference --use-bvh")
    return run, blocks, bpp, cores, how


def run_reference(args, rank, wl):
    """The reference's algorithm on the host CPU (oracle port, all cores)."""
    if rank != 0:
        return
    label, kind, width, height = wl[:4]
    base = {"impl": "reference", "metric": "Mrays/s (all bounces)", "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    run, blocks, bpp, cores, how = _oracle_setup(wl)
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
        "config": {"workload": label, "note": f"CPU restatement of reference GLSL ({how}), {cores} host threads — the "
                   "reference's wgpu/lavapipe path cannot run in this image"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    emit(base)


def cpu_baseline_sample(wl):
    """Bounded CPU sample for the main line (rank 0, N = 1): ~20 s of oracle work."""
    label, kind, width, height = wl[:4]
    run, blocks, bpp, cores, how = _oracle_setup(wl)
    mid = bpp // 2
    _, dt = run(blocks[mid:mid + 2])
    per_block = max(dt / 2, 1e-4)
    nb = int(min(len(blocks), max(2, 20.0 / per_block)))
    rays, dt = run(blocks[:nb])
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"first {nb} ImageBlocks ({nb / bpp:.1f} sample passes of {width}x{height}) of the job, "
                      f"{rays / 1e6:.1f} Mrays in {dt:.1f} s ({how})"}


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


class Rig:
    """One rank's device state: the context on this rank's GPU, running on a torch stream, joined to the other ranks
    through the LIBRARY's communicator (hjk_comm_unique_id / hjk_comm_init; the 128-byte id travels over the torchrun
    rendezvous).  torch.distributed carries only the barrier and the gathering of the timings."""

    def __init__(self, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        import hijiki_b200 as hj
        self.torch, self.dist, self.hj = torch, dist, hj
        self.rank, self.world, self.local_rank = rank, world, local_rank
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
        torch.cuda.set_device(local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        self.ctx = hj.Context(local_rank)
        self.stream = torch.cuda.Stream(device=local_rank)
        self.ctx.set_stream(self.stream.cuda_stream)
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                uid = torch.frombuffer(bytearray(hj.comm_unique_id()), dtype=torch.uint8).clone()
            uid = uid.cuda()
            dist.broadcast(uid, 0)
            self.ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        if self.world == 1:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(self, values):
        if self.world == 1:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t)
        return t.tolist()

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)


def run_workload(rig, wl, steps, warmup, sps, want_e2e, want_exact, sample_clocks):
    """Times one workload on every rank.  Returns the merged (max time, summed rays) record on every rank."""
    hj, torch = rig.hj, rig.torch
    ctx, stream, rank, world = rig.ctx, rig.stream, rig.rank, rig.world
    label, kind, width, height, spp, max_bounces, default_sps = wl
    sps = sps or default_sps
    n_steps_total = warmup + steps
    gen = hj.ImageBlockGenerator(width, height, BLOCK, spp)
    bpp = gen.blocks_per_pass
    # only the passes the run touches are generated (the list is a prefix of the job's list)
    gen.num_samples = min(spp, n_steps_total * sps * world)
    blocks = gen.blocks()
    t0 = time.perf_counter()
    compiled = make_scene(kind).compile(use_bvh=False)
    t_scene = time.perf_counter() - t0
    rig.barrier()
    t0 = time.perf_counter()
    ctx.scene_upload(compiled)  # N > 1: rank 0 builds the wide BVH, ncclBroadcast carries it to the others
    ctx.synchronize()
    t_upload = time.perf_counter() - t0
    ctx.frame_begin(width, height)
    ctx.set_profiling(True)
    params = hj.make_params(max_bounces=max_bounces)

    def reduce():
        if world > 1:
            ctx.reduce_frame(0, timed=False)  # one ncclReduce(sum) of the W x H float4 frame to rank 0

    # ---------------- device-resident arm: block lists uploaded before the timed region
    slices = [step_slice(blocks, bpp, s, rank, world, sps) for s in range(n_steps_total)]
    handles = [ctx.blocks_upload(s) for s in slices]
    for s in range(warmup):
        ctx.frame_begin(width, height)
        ctx.render_resident(handles[s], 0, slices[s].size, params)
        reduce()
    rig.barrier()
    sampler = ClockSampler(rig.local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
    ev0, ev1 = rig.event(), rig.event()
    rays = ext_rays = sh_rays = launches = 0
    kernel_ms = {}
    ev0.record(stream)
    for s in range(warmup, n_steps_total):
        ctx.frame_begin(width, height)
        st = ctx.render_resident(handles[s], 0, slices[s].size, params)
        reduce()
        rays += st.n_rays
        ext_rays += st.n_extension_rays
        sh_rays += st.n_shadow_rays
        launches += st.n_launches + (1 if world > 1 else 0)
        for k, v in st.kernel_ms.items():
            kernel_ms[k] = kernel_ms.get(k, 0.0) + v
    ev1.record(stream)
    rig.barrier()
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1)

    # exact-tie mode (HJK_RENDER_EXACT_TIES), two steps: what bit-identity with the reference on every ray costs
    exact = None
    if want_exact:
        exact_params = hj.make_params(max_bounces=max_bounces, flags=hj.HJK_RENDER_EXACT_TIES)
        exact_rays, exact_ms = 0, 0.0
        for s in range(warmup, min(n_steps_total, warmup + 2)):
            ctx.frame_begin(width, height)
            st = ctx.render_resident(handles[s], 0, slices[s].size, exact_params)
            exact_rays += st.n_rays
            exact_ms += st.ms_total
        exact = {"value": exact_rays / (exact_ms * 1e-3) / 1e6 if exact_ms else None, "unit": "Mrays/s",
                 "unresolved_clusters_last_step": ctx.get_info("unresolved_ties"),
                 "note": "this rank, 2 steps, HJK_RENDER_EXACT_TIES (the mode whose frames the parity tests hold "
                         "bit-identical to the oracle at these sizes, tests/test_gpu_fullsize.py); unresolved = tie clusters "
                         "longer than the 8*M_EPS window or 12 candidates, counted, each can move one sample"}
    for h in handles:
        ctx.blocks_free(h)

    # ---------------- end-to-end arm: host block list in, normalised frame out (on rank 0), every step
    e2e_ms, e2e_rays, result_mean = None, 0, None
    if want_e2e:
        pinned_blocks = [torch.from_numpy(s.view(np.uint8).copy()).pin_memory() for s in slices]
        # two host frames: the copy of step k (hjk_readback_begin, on the library's copy stream) overlaps the
        # rendering of step k + 1; every step's frame is read back inside the timed region, the last one awaited in it
        out_host = [torch.empty((height, width, 4), dtype=torch.float32).pin_memory() for _ in range(2)] if rank == 0 else None

        def e2e_step(s):
            ctx.frame_begin(width, height)
            st = ctx.render((pinned_blocks[s].data_ptr(), slices[s].size), params)
            ctx.readback_begin_ptr(0, out_host[s & 1].data_ptr() if rank == 0 else 0, width * 16, normalise=True)
            return st.n_rays

        for s in range(warmup):
            e2e_step(s)
        ctx.readback_wait()
        rig.barrier()
        e0, e1 = rig.event(), rig.event()
        e0.record(stream)
        for s in range(warmup, n_steps_total):
            e2e_rays += e2e_step(s)
        ctx.readback_wait()
        e1.record(stream)
        rig.barrier()
        e2e_ms = e0.elapsed_time(e1)
        if rank == 0:
            result_mean = float(out_host[(n_steps_total - 1) & 1][..., :3].mean())

    # the path's one collective, timed alone
    reduce_info = None
    if world > 1:
        rig.barrier()
        a0, a1 = rig.event(), rig.event()
        reps = 20
        out = {}
        for name, root in (("reduce_to_root", 0), ("allreduce", -1)):
            ctx.reduce_frame(root, timed=False)
            rig.barrier()
            a0.record(stream)
            for _ in range(reps):
                ctx.reduce_frame(root, timed=False)
            a1.record(stream)
            rig.barrier()
            out[name] = a0.elapsed_time(a1) / reps
        nbytes = width * height * 16
        ar = max(out["allreduce"], 1e-6)
        reduce_info = {"bytes": nbytes, "reduce_to_root_ms": out["reduce_to_root"], "allreduce_ms": ar,
                       "allreduce_bus_gbs": 2 * (world - 1) / world * nbytes / (ar * 1e-3) / 1e9,
                       "note": "hjk_reduce_frame on the render stream (the library's communicator, NCCL by dlopen), 20 "
                               "back-to-back calls; bus bandwidth = 2(N-1)/N * S / t (nominal NVLink 5: 900 GB/s per "
                               "direction)"}

    info = {k: ctx.get_info(k) for k in ("bvh_nodes", "bvh_prims", "bvh_depth", "bvh_bytes", "wave_paths",
                                          "blocks_per_sm_traverse", "blocks_per_sm_tile", "n_sms", "stack_overflows")}

    ms, e2e_max = rig.max_over_ranks([ms, e2e_ms or 0.0])
    rays, e2e_rays, ext_rays, sh_rays, launches = (int(v) for v in
                                                   rig.sum_over_ranks([rays, e2e_rays, ext_rays, sh_rays, launches]))
    peak, peak_src = measured_peaks()
    wave_passes = max(1, min(sps, (info["wave_paths"] + width * height // 2) // (width * height)))
    n_ext_launches = steps * (max_bounces + 1) * -(-sps // wave_passes)
    ext_ms = kernel_ms.get("extend", 0.0)  # slot HJK_K_EXTEND times k_trace (extension + shadow rays) on this rank
    trace_bytes = (EXTEND_BYTES_PER_RAY * ext_rays + SHADOW_BYTES_PER_RAY * sh_rays) / world
    scene_bytes = SCENE_FETCH_BYTES_PER_RAY * rays / world if kind == "terrain" else 0.0
    achieved = (trace_bytes + scene_bytes) / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else None
    pipe = (PIPE_BYTES_EXT * ext_rays + PIPE_BYTES_SHADOW * sh_rays + scene_bytes * world) / world / (ms * 1e-3) / 1e9
    nm = ncu_metrics(kind)
    rec = {
        "value": rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps,
        "config": {"workload": label, "spp_per_step_per_gpu": sps, "blocks_per_step_per_gpu": sps * bpp,
                   "parallelism": f"sample-pass dp{world}",
                   "l2": "inputs larger than L2: the path state + queues of one wave (~200 B per path, GBs per wave) "
                         "exceed the 126 MB L2",
                   "image_mean": result_mean, "bvh": info,
                   "scene_build_s": round(t_scene, 2), "bvh_build_upload_s": round(t_upload, 2)},
        "rays": {"extension": ext_rays, "shadow": sh_rays, "per_path": rays / max(1, steps * sps * width * height * world)},
        "gpu_launches": launches,
        "kernel_ms_per_step": {k: v / steps for k, v in kernel_ms.items()},
        "roofline": {"kernel": "k_trace_coop", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None,
                     "traffic": (nm["dram_bytes_per_ray"] * rays / world / n_ext_launches
                                 if nm and nm.get("dram_bytes_per_ray") and n_ext_launches else None),
                     "algorithmic_bytes_per_launch": (trace_bytes + scene_bytes) / n_ext_launches if n_ext_launches else None,
                     "peak_source": peak_src,
                     # what actually bounds traversal on a cache-resident scene (SURVEY 8d caveat): cycles with an
                     # instruction issued / lanes per instruction, from the committed ncu summary (not measured live)
                     "ncu": nm,
                     "bytes_per_ray": {"extension": EXTEND_BYTES_PER_RAY, "shadow": SHADOW_BYTES_PER_RAY,
                                       "scene_fetch": SCENE_FETCH_BYTES_PER_RAY if kind == "terrain" else 0},
                     "launches": n_ext_launches,
                     "avg_launch_ms": ext_ms / n_ext_launches if n_ext_launches else None,
                     "share_of_step": ext_ms / ms if ms else None,
                     "note": "traversal of a cache-resident BVH is issue/latency-bound, not HBM-bound "
                             "(SURVEY.md §8d); see profiles/ for SM issue utilisation"},
        "pipeline_bytes": {"achieved": pipe, "unit": "GB/s", "frac": pipe / peak},
    }
    if clocks is not None:
        rec["clocks"] = clocks
    if exact is not None:
        rec["exact_ties"] = exact
    if e2e_ms is not None:
        rec["e2e"] = {"value": e2e_rays / (e2e_max * 1e-3) / 1e6, "unit": "Mrays/s",
                      "h2d_bytes_per_step": int(slices[0].nbytes) * world, "d2h_bytes_per_step": width * height * 16,
                      "ms_per_step": e2e_max / steps,
                      "note": "bytes are whole-job per step: every rank uploads its block list, rank 0 alone reads the frame back"}
    if reduce_info:
        rec["reduce"] = reduce_info
    return rec


def run_strong(rig, wl, total_spp):
    """A fixed job split over the ranks: `total_spp` sample passes of the workload, pass p -> rank p mod N, from the
    scene upload (rank 0 builds the wide BVH and broadcasts it) to the normalised frame in rank 0's host memory.
    Wall clock between two barriers, max over ranks — the one number here that is not device-timed, because the
    BVH build is host work; the device time of the render calls is reported beside it."""
    hj, torch = rig.hj, rig.torch
    ctx, rank, world = rig.ctx, rig.rank, rig.world
    label, kind, width, height, spp, max_bounces, _ = wl
    gen = hj.ImageBlockGenerator(width, height, BLOCK, total_spp)
    mine = hj.split_passes(gen.blocks(), gen.blocks_per_pass, rank, world)
    compiled = make_scene(kind).compile(use_bvh=False)
    out_host = torch.empty((height, width, 4), dtype=torch.float32).pin_memory() if rank == 0 else None
    params = hj.make_params(max_bounces=max_bounces)
    res = []
    for attempt in range(2):  # the second run is reported (first: allocations, NCCL channel setup)
        rig.barrier()
        t0 = time.perf_counter()
        ctx.scene_upload(compiled)
        t1 = time.perf_counter()
        ctx.frame_begin(width, height)
        st = ctx.render(mine, params) if len(mine) else None
        ctx.readback_root_ptr(0, out_host.data_ptr() if rank == 0 else 0, width * 16, normalise=True)
        rig.barrier()
        t2 = time.perf_counter()
        res = [t2 - t0, t1 - t0, st.ms_total if st else 0.0, float(st.n_rays) if st else 0.0]
    job_s, upload_s, render_ms = rig.max_over_ranks(res[:3])
    rays = rig.sum_over_ranks([res[3]])[0]
    return {"workload": f"{label.split(',')[0]} {width}x{height}, {total_spp} spp in total, max {max_bounces} bounces",
            "scaling": "strong", "n_gpus": world, "job_s": job_s, "scene_upload_s": upload_s,
            "render_device_ms": render_ms, "value": rays / job_s / 1e6, "unit": "Mrays/s",
            "render_only_value": rays / (render_ms * 1e-3) / 1e6 if render_ms else None,
            "note": "wall clock from hjk_scene_upload to the frame on rank 0's host (BVH build + broadcast, render, "
                    "ncclReduce, readback), max over ranks"}


def run_denoiser(rig, params):
    """Reconstruction ("denoiser") bandwidth on 3840x2160 feature buffers (configs[4]), per GPU."""
    hj = rig.hj
    dw, dh = 3840, 2160
    rng = np.random.default_rng(5 + rig.rank)
    rad = np.exp(rng.standard_normal((dh, dw, 4), dtype=np.float32))
    rad[..., 3] = 1.0
    nrm = rng.standard_normal((dh, dw, 4), dtype=np.float32)
    nrm[..., :3] /= np.linalg.norm(nrm[..., :3], axis=2, keepdims=True)
    dblocks = hj.ImageBlockGenerator(dw, dh, BLOCK, 1).blocks()
    dctx = hj.Context(rig.local_rank)
    dctx.frame_begin(dw, dh)
    dctx.denoise_upload(rad, nrm, dblocks)
    dctx.denoise_resident(params, 5)
    reps = 50
    dms = dctx.denoise_resident(params, reps)
    dctx.close()
    dms = rig.max_over_ranks([dms])[0]
    gbs = RECON_BYTES_PER_PX * dw * dh * reps / (dms * 1e-3) / 1e9
    peak, _ = measured_peaks()
    return {"workload": "3840x2160 synthetic feature buffers (random unit normals), 510 ImageBlocks/pass, R=2 "
                        "(BASELINE.json configs[4]), per GPU (max time over ranks)",
            "kernel": "k_recon", "ms_per_pass": dms / reps, "bytes_per_px": RECON_BYTES_PER_PX, "achieved": gbs,
            "unit": "GB/s", "peak": peak, "frac": gbs / peak, "aggregate_gbs": gbs * rig.world}


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
    ap.add_argument("--no-extras", action="store_true", help="skip the terrain / spheres / spheres64 / strong objects")
    ap.add_argument("--cpu-seconds", type=float, default=100.0, help="wall budget of the --impl reference run")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, wl)
        return
    args.warmup = max(args.warmup, 3)

    rig = Rig(rank, world, local_rank)
    main_rec = run_workload(rig, wl, args.steps, args.warmup, args.spp_per_step, not args.no_e2e, True, True)
    line = {"metric": "Mrays/s (all bounces)", "value": main_rec.pop("value"), "unit": main_rec.pop("unit"),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_rec.pop("ms_per_step"),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    main_rec.pop("steps"), main_rec.pop("warmup")
    line.update(main_rec)
    if not args.no_extras and args.workload == "cbox1080":
        # the other BASELINE configs, measured the same way (device-resident + end to end), fewer steps
        for name in ("terrain", "spheres", "spheres64"):
            rec = run_workload(rig, WORKLOADS[name], min(args.steps, 3), 3, 0, not args.no_e2e, name != "spheres64", False)
            rec["metric"] = "Mrays/s (all bounces)"
            line[name] = rec
        line["strong"] = run_strong(rig, wl, 256)
    params = rig.hj.make_params(max_bounces=wl[5])
    rig.ctx.close()
    if not args.no_denoiser:
        line["denoiser"] = run_denoiser(rig, params)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample(wl)
        emit(line)
    if world > 1:
        rig.dist.barrier()
        rig.dist.destroy_process_group()


if __name__ == "__main__":
    main()
