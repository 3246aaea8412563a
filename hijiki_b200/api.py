"""Host-side mirror of the reference's Rust front-end for the path-tracing hot path.

Names follow reference ``src/main.rs``:

* ``Scene.from_obj`` / ``Scene.compile``          src/main.rs:413-530, 172-358
* ``ImageBlockGenerator``                          src/main.rs:619-682
* ``Renderer.new / render / save_image``           src/main.rs:1167-1423

Everything that touches pixels runs in the CUDA library behind ``include/hijiki_b200.h``;
this module only marshals pointers.  There is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _abi
from ._abi import (BLOCK_DTYPE, RAY_DTYPE, HijikiError, HjkParams, HjkScene, HjkStats, as_ptr)

DEFAULT_ROOT_SEED = 0x48494A494B49  # recorded splitmix64 stream replacing OS entropy


def _check_host(lib, rc: int) -> None:
    if rc != 0:
        raise HijikiError(rc, lib.hjk_host_last_error().decode())


class CompiledScene:
    """The 12 ``CompiledScene`` arrays (reference src/main.rs:376-397) as numpy views."""

    def __init__(self, lib, handle):
        self._lib, self._handle = lib, handle
        self.view = HjkScene()
        _check_host(lib, lib.hjk_host_scene_view(handle, C.byref(self.view)))

    def array(self, name: str) -> np.ndarray:
        arr = getattr(self.view, name)
        size, dtype, per = _abi.SCENE_ELEM[name]
        if arr.count == 0 or not arr.ptr:
            return np.zeros((0, per), dtype=dtype)
        buf = (C.c_uint8 * (arr.count * size)).from_address(arr.ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(arr.count, per)

    @property
    def info(self) -> _abi.HjkSceneInfo:
        return _abi.HjkSceneInfo.from_address(self.view.scene.ptr)

    def bvh_stats(self, pad_rel: float = -1.0) -> dict:
        out = np.zeros(6, dtype=np.uint64)
        _check_host(self._lib, self._lib.hjk_host_bvh_stats(C.byref(self.view), pad_rel, as_ptr(out)))
        return {"nodes": int(out[0]), "prims": int(out[1]), "depth": int(out[2]), "valid": bool(out[3]),
                "sah_cost": out[4] / 1000.0, "bytes": int(out[5])}

    def bvh_digest(self, n_threads: int = 0, pad_rel: float = -1.0) -> int:
        """FNV-1a of the wide BVH built on at most ``n_threads`` host threads (0 = all)."""
        out = np.zeros(1, dtype=np.uint64)
        _check_host(self._lib, self._lib.hjk_host_bvh_digest(C.byref(self.view), pad_rel, n_threads, as_ptr(out)))
        return int(out[0])

    def close(self) -> None:
        if self._handle:
            self._lib.hjk_host_scene_free(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """``Scene`` of the reference (src/main.rs:162-170); compiled eagerly by the C++ host."""

    def __init__(self, kind: str, **kw):
        self.kind, self.kw = kind, kw

    @classmethod
    def from_obj(cls, path: str, put_cbox_spheres: bool = False) -> "Scene":
        """``Scene::from_obj`` (+ main()'s ``--put-cbox-spheres`` block, src/main.rs:1463-1483)."""
        return cls("obj", path=os.fspath(path), put_cbox_spheres=put_cbox_spheres)

    @classmethod
    def terrain(cls, grid_n: int = 2237, seed: int = 7) -> "Scene":
        """BASELINE.json config 3: 2*grid_n^2-triangle checkerboard terrain + emissive quads."""
        return cls("terrain", grid_n=grid_n, seed=seed)

    @classmethod
    def spheres(cls, lattice_n: int = 8, seed: int = 7) -> "Scene":
        """BASELINE.json config 4: lattice_n^3 dielectric/mirror spheres."""
        return cls("spheres", lattice_n=lattice_n, seed=seed)

    def compile(self, use_bvh: bool = False) -> CompiledScene:
        """``Scene::compile``.  ``use_bvh`` also emits the reference-layout ``bvh`` binding."""
        lib = _abi.load()
        h = C.c_void_p()
        if self.kind == "obj":
            rc = lib.hjk_host_scene_from_obj(self.kw["path"].encode(), int(self.kw["put_cbox_spheres"]),
                                             int(use_bvh), C.byref(h))
        elif self.kind == "terrain":
            rc = lib.hjk_host_scene_terrain(self.kw["grid_n"], self.kw["seed"], int(use_bvh), C.byref(h))
        else:
            rc = lib.hjk_host_scene_spheres(self.kw["lattice_n"], self.kw["seed"], int(use_bvh), C.byref(h))
        _check_host(lib, rc)
        return CompiledScene(lib, h)


class ImageBlockGenerator:
    """``ImageBlockGenerator`` (src/main.rs:619-682) with ``rand::random()`` replaced by a
    recorded splitmix64 stream (``root_seed``), so that a run can be repeated and checked."""

    def __init__(self, width: int, height: int, block_size: int = 128, num_samples: int = 64,
                 root_seed: int = DEFAULT_ROOT_SEED):
        if block_size & 63:
            raise ValueError("block_size must be a multiple of 64 (src/main.rs:633)")
        self.width, self.height, self.block_size = width, height, block_size
        self.num_samples, self.root_seed = num_samples, root_seed

    def blocks(self) -> np.ndarray:
        lib = _abi.load()
        n = lib.hjk_host_generate_blocks(self.width, self.height, self.block_size, self.num_samples,
                                         self.root_seed, None, 0)
        out = np.zeros(n, dtype=BLOCK_DTYPE)
        if n:
            lib.hjk_host_generate_blocks(self.width, self.height, self.block_size, self.num_samples,
                                         self.root_seed, as_ptr(out), n)
        return out

    def __iter__(self):
        return iter(self.blocks())

    @property
    def blocks_per_pass(self) -> int:
        bs = self.block_size
        return ((self.width + bs - 1) // bs) * ((self.height + bs - 1) // bs)


def split_passes(blocks: np.ndarray, blocks_per_pass: int, rank: int, world: int) -> np.ndarray:
    """Sample-pass data parallelism: pass p goes to rank p mod world."""
    if world == 1:
        return blocks
    n_pass = len(blocks) // blocks_per_pass
    keep = [p for p in range(n_pass) if p % world == rank]
    idx = np.concatenate([np.arange(p * blocks_per_pass, (p + 1) * blocks_per_pass) for p in keep]) if keep \
        else np.zeros(0, dtype=np.int64)
    return np.ascontiguousarray(blocks[idx])


@dataclass
class RenderStats:
    n_paths: int
    n_extension_rays: int
    n_shadow_rays: int
    ms_total: float
    kernel_ms: dict
    n_launches: int

    @property
    def n_rays(self) -> int:
        return self.n_extension_rays + self.n_shadow_rays

    @property
    def mrays_per_s(self) -> float:
        return self.n_rays / (self.ms_total * 1e-3) / 1e6 if self.ms_total > 0 else 0.0


def make_params(max_bounces: int = 1000, rr_start: int = 3, recon_radius: int = 2, recon_stddev: float = 0.5,
                eps: float = 1e-4, flags: int = 0) -> HjkParams:
    """Constants the reference hard-codes (render.glsl:92,137; src/main.rs:1284-1285; math.glsl:2)."""
    return HjkParams(max_bounces, rr_start, recon_radius, recon_stddev, eps, flags)


class Context:
    """Owner of one GPU's device state (``GPU::new``, src/main.rs:692-712)."""

    def __init__(self, device=0):
        """``device``: one CUDA device id, or a sequence of ids for a single-process multi-GPU group."""
        self.lib = _abi.load()
        self.ptr = C.c_void_p()
        ids = [int(device)] if isinstance(device, (int, np.integer)) else [int(d) for d in device]
        dev = (C.c_int * len(ids))(*ids)
        rc = self.lib.hjk_create(dev, len(ids), C.byref(self.ptr))
        if rc != 0:
            raise HijikiError(rc, self.lib.hjk_last_error(None).decode())
        self._keep = []

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise HijikiError(rc, self.lib.hjk_last_error(self.ptr).decode())

    def close(self) -> None:
        if self.ptr:
            self.lib.hjk_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene / frame
    def scene_upload(self, compiled: CompiledScene) -> None:
        self._check(self.lib.hjk_scene_upload(self.ptr, C.byref(compiled.view)))

    def frame_begin(self, width: int, height: int) -> None:
        self._check(self.lib.hjk_frame_begin(self.ptr, width, height))

    # ---- render
    @staticmethod
    def _stats(st: HjkStats) -> RenderStats:
        return RenderStats(st.n_paths, st.n_extension_rays, st.n_shadow_rays, st.ms_total,
                           {k: st.kernel_ms[i] for i, k in enumerate(_abi.KERNEL_SLOTS) if not k.startswith("_")},
                           st.n_launches)

    def render(self, blocks, params: HjkParams, want_stats: bool = True) -> RenderStats | None:
        """``Renderer::render`` over a HOST block list (numpy BLOCK_DTYPE array or raw pointer+count)."""
        st = HjkStats()
        if isinstance(blocks, tuple):
            p, n = blocks
        else:
            blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
            p, n = blocks.ctypes.data, blocks.size
        self._check(self.lib.hjk_render(self.ptr, C.c_void_p(p), n, C.byref(params),
                                        C.byref(st) if want_stats else None))
        return self._stats(st) if want_stats else None

    def blocks_upload(self, blocks: np.ndarray) -> int:
        blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        h = C.c_uint64()
        self._check(self.lib.hjk_blocks_upload(self.ptr, as_ptr(blocks), blocks.size, C.byref(h)))
        return h.value

    def render_resident(self, handle: int, first: int, count: int, params: HjkParams,
                        want_stats: bool = True) -> RenderStats | None:
        st = HjkStats()
        self._check(self.lib.hjk_render_resident(self.ptr, handle, first, count, C.byref(params),
                                                 C.byref(st) if want_stats else None))
        return self._stats(st) if want_stats else None

    def blocks_free(self, handle: int) -> None:
        self._check(self.lib.hjk_blocks_free(self.ptr, handle))

    # ---- results
    def readback(self, normalise: bool = True, out: np.ndarray | None = None) -> np.ndarray:
        w, h = self.frame_size()
        if out is None:
            out = np.empty((h, w, 4), dtype=np.float32)
        self._check(self.lib.hjk_readback(self.ptr, as_ptr(out), out.strides[0], int(normalise)))
        return out

    def readback_ptr(self, ptr: int, pitch: int, normalise: bool = True) -> None:
        self._check(self.lib.hjk_readback(self.ptr, C.c_void_p(ptr), pitch, int(normalise)))

    def readback_root(self, root: int, normalise: bool = True, out: np.ndarray | None = None):
        """Frame summed over the ranks and delivered to ``root`` only; the other ranks get ``None``."""
        if self.get_info("n_ranks") > 1 and self.get_info("n_devices") == 1 and 0 <= root != self.get_info("rank"):
            self._check(self.lib.hjk_readback_root(self.ptr, root, None, 0, int(normalise)))
            return None
        w, h = self.frame_size()
        if out is None:
            out = np.empty((h, w, 4), dtype=np.float32)
        self._check(self.lib.hjk_readback_root(self.ptr, root, as_ptr(out), out.strides[0], int(normalise)))
        return out

    def readback_root_ptr(self, root: int, ptr: int, pitch: int, normalise: bool = True) -> None:
        self._check(self.lib.hjk_readback_root(self.ptr, root, C.c_void_p(ptr) if ptr else None, pitch, int(normalise)))

    def readback_begin_ptr(self, root: int, ptr: int, pitch: int, normalise: bool = True) -> None:
        """``readback_root_ptr`` without waiting for the copy (``readback_wait`` completes it)."""
        self._check(self.lib.hjk_readback_begin(self.ptr, root, C.c_void_p(ptr) if ptr else None, pitch, int(normalise)))

    def readback_wait(self) -> None:
        self._check(self.lib.hjk_readback_wait(self.ptr))

    def read_features(self, root: int = -1):
        """Averaged first-hit (normal, depth) of the frame (option ``feature_buffers``), reduced like the frame."""
        if self.get_info("n_ranks") > 1 and self.get_info("n_devices") == 1 and 0 <= root != self.get_info("rank"):
            self._check(self.lib.hjk_read_features(self.ptr, root, None, 0))
            return None
        w, h = self.frame_size()
        out = np.empty((h, w, 4), dtype=np.float32)
        self._check(self.lib.hjk_read_features(self.ptr, root, as_ptr(out), out.strides[0]))
        return out

    def reduce_frame(self, root: int = -1, timed: bool = True) -> float:
        """Sum of the frame over the ranks (to every rank, or to ``root``).  ``timed`` waits for the collective and
        returns its device time; otherwise the call only enqueues it."""
        if not timed:
            self._check(self.lib.hjk_reduce_frame(self.ptr, root, None))
            return 0.0
        ms = C.c_float()
        self._check(self.lib.hjk_reduce_frame(self.ptr, root, C.byref(ms)))
        return ms.value

    def read_intermediate(self, layer: int) -> np.ndarray:
        w, h = self.frame_size()
        out = np.empty((h, w, 4), dtype=np.float32)
        self._check(self.lib.hjk_read_intermediate(self.ptr, layer, as_ptr(out)))
        return out

    def frame_size(self):
        return self.get_info("width"), self.get_info("height")

    def trace_first_hit(self, rays: np.ndarray, any_hit: bool = False, exact_ties: bool = False,
                        eps: float | None = None):
        """Parity hook: closest hit (or occlusion) of scene.glsl:97-175 on a ray batch."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        n = rays.size
        ids = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float32)
        uv = np.empty((n, 2), dtype=np.float32)
        mode = int(any_hit) | (2 if exact_ties else 0)
        if eps is None:
            self._check(self.lib.hjk_trace_first_hit(self.ptr, as_ptr(rays), n, mode, as_ptr(ids), as_ptr(t), as_ptr(uv)))
        else:
            self._check(self.lib.hjk_trace_first_hit_eps(self.ptr, as_ptr(rays), n, mode, eps, as_ptr(ids), as_ptr(t),
                                                         as_ptr(uv)))
        return ids, t, uv

    def denoise_pass(self, radiance, normal_depth, albedo, blocks, params: HjkParams) -> None:
        """``ReconstructionPipeline::run`` (src/main.rs:992-1003) on caller-supplied layers."""
        radiance = np.ascontiguousarray(radiance, dtype=np.float32)
        normal_depth = np.ascontiguousarray(normal_depth, dtype=np.float32)
        if albedo is not None:
            albedo = np.ascontiguousarray(albedo, dtype=np.float32)
        blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        self._check(self.lib.hjk_denoise_pass(self.ptr, as_ptr(radiance), as_ptr(normal_depth),
                                              as_ptr(albedo) if albedo is not None else None, as_ptr(blocks),
                                              blocks.size, C.byref(params)))

    def denoise_upload(self, radiance, normal_depth, blocks) -> None:
        radiance = np.ascontiguousarray(radiance, dtype=np.float32)
        normal_depth = np.ascontiguousarray(normal_depth, dtype=np.float32)
        blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        self._check(self.lib.hjk_denoise_upload(self.ptr, as_ptr(radiance), as_ptr(normal_depth), as_ptr(blocks),
                                                blocks.size))

    def denoise_resident(self, params: HjkParams, repeat: int = 1) -> float:
        ms = C.c_float()
        self._check(self.lib.hjk_denoise_resident(self.ptr, C.byref(params), repeat, C.byref(ms)))
        return ms.value

    # ---- plumbing
    def accumulator_device_ptr(self):
        p, n = C.c_uint64(), C.c_uint64()
        self._check(self.lib.hjk_accumulator_device_ptr(self.ptr, C.byref(p), C.byref(n)))
        return p.value, n.value

    def synchronize(self) -> None:
        self._check(self.lib.hjk_synchronize(self.ptr))

    def set_stream(self, cuda_stream: int) -> None:
        self._check(self.lib.hjk_set_stream(self.ptr, C.c_void_p(cuda_stream)))

    def set_profiling(self, on: bool) -> None:
        self._check(self.lib.hjk_set_profiling(self.ptr, int(on)))

    def set_option(self, key: str, value: int) -> None:
        self._check(self.lib.hjk_set_option(self.ptr, key.encode(), int(value)))

    def get_info(self, key: str) -> int:
        v = C.c_int64()
        self._check(self.lib.hjk_get_info(self.ptr, key.encode(), C.byref(v)))
        return v.value

    def comm_init(self, id128: bytes, rank: int, world: int) -> None:
        buf = C.create_string_buffer(id128, 128)
        self._check(self.lib.hjk_comm_init(self.ptr, buf, rank, world))

    def allreduce_accumulator(self) -> float:
        ms = C.c_float()
        self._check(self.lib.hjk_allreduce_accumulator(self.ptr, C.byref(ms)))
        return ms.value


def comm_unique_id() -> bytes:
    lib = _abi.load()
    buf = C.create_string_buffer(128)
    rc = lib.hjk_comm_unique_id(buf)
    if rc != 0:
        raise HijikiError(rc, lib.hjk_last_error(None).decode())
    return buf.raw


class Renderer:
    """``Renderer`` of the reference (src/main.rs:1143-1424): ``new`` uploads the scene and
    zeroes the accumulator, ``render`` integrates + reconstructs every block, ``save_image``
    reads back, divides by the weight and writes a 3-channel float EXR."""

    def __init__(self, scene: Scene, generator: ImageBlockGenerator, present_interval: int = 128,
                 use_bvh: bool = False, device=0, max_bounces: int = 1000, rank: int = 0, world: int = 1,
                 comm_id: bytes | None = None):
        # present_interval drove the reference's preview window (src/main.rs:1335-1340): no-op here.
        # use_bvh selected scene.glsl's USE_BVH walk; this path always walks its own wide BVH.
        del present_interval
        self.generator = generator
        self.compiled = scene.compile(use_bvh=use_bvh)
        self.ctx = Context(device)
        if comm_id is not None and world > 1:  # before the upload: rank 0 builds the BVH, the others receive it
            self.ctx.comm_init(comm_id, rank, world)
        self.ctx.scene_upload(self.compiled)
        self.ctx.frame_begin(generator.width, generator.height)
        self.params = make_params(max_bounces=max_bounces)
        self.blocks = split_passes(generator.blocks(), generator.blocks_per_pass, rank, world)
        self.stats: RenderStats | None = None

    @classmethod
    def new(cls, scene, generator, present_interval=128, use_bvh=False, **kw) -> "Renderer":
        return cls(scene, generator, present_interval, use_bvh, **kw)

    def render(self) -> RenderStats:
        if len(self.blocks) == 0:  # more ranks than sample passes: this rank keeps its zeroed frame and still
            self.stats = RenderStats(0, 0, 0, 0.0, {}, 0)  # joins the frame's collective in image()
            return self.stats
        self.stats = self.ctx.render(self.blocks, self.params)
        return self.stats

    def image(self) -> np.ndarray:
        """(H, W, 4) float32: rgb / weight, weight (src/main.rs:1399)."""
        return self.ctx.readback(normalise=True)

    def save_image(self, path: str) -> None:
        img = self.image()
        lib = self.ctx.lib
        rc = lib.hjk_host_write_exr(os.fspath(path).encode(), as_ptr(img), img.shape[1], img.shape[0], img.strides[0])
        _check_host(lib, rc)
