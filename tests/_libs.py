"""ctypes access to the two pieces of TEST infrastructure:

* ``oracle/liboracle.so``           — CPU restatement of the reference shaders (the checker);
* ``tests/native/libhjk_hosttest.so`` — the product's device headers compiled for the host, so
  kernel logic can be checked in the CPU-only container.

Neither is ever loaded by the product package ``hijiki_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
NATIVE_DIR = os.path.join(ROOT, "tests", "native")
CBOX_OBJ = os.path.join(ROOT, "scenes", "cbox", "cbox.obj")

import sys  # noqa: E402

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hijiki_b200 import _abi  # noqa: E402  (struct layouts only; does not load the CUDA library)

_P = C.c_void_p


def _make(directory: str) -> None:
    subprocess.run(["make", "-s", "-C", directory], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


class OrcParams(C.Structure):
    _fields_ = [("max_bounces", C.c_uint32), ("rr_start", C.c_uint32), ("recon_radius", C.c_uint32),
                ("recon_stddev", C.c_float), ("eps", C.c_float), ("use_bvh", C.c_uint32),
                ("block_size", C.c_uint32), ("skip_recon", C.c_uint32)]


class OrcStats(C.Structure):
    _fields_ = [("n_paths", C.c_uint64), ("n_extension_rays", C.c_uint64), ("n_shadow_rays", C.c_uint64),
                ("seconds", C.c_double)]


class OrcPathVertex(C.Structure):
    _fields_ = [("shape_id", C.c_int32), ("t", C.c_float), ("rng_after", C.c_uint32),
                ("throughput", C.c_float * 3), ("total", C.c_float * 3), ("shadow_state", C.c_int32),
                ("dir_len2", C.c_float)]


_oracle = None
_hosttest = None


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        _make(ORACLE_DIR)
        L = C.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))
        L.orc_seed_rng.restype = C.c_uint32
        L.orc_seed_rng.argtypes = [C.c_uint32]
        L.orc_rand_uint.restype = C.c_uint32
        L.orc_rand_uint.argtypes = [C.POINTER(C.c_uint32)]
        L.orc_rand_uniform_float.restype = C.c_float
        L.orc_rand_uniform_float.argtypes = [C.POINTER(C.c_uint32)]
        L.orc_camera_ray.restype = None
        L.orc_camera_ray.argtypes = [_P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _P]
        L.orc_recon_spatial_weights.restype = None
        L.orc_recon_spatial_weights.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float, _P]
        L.orc_trace.restype = C.c_int
        L.orc_trace.argtypes = [C.POINTER(_abi.HjkScene), _P, C.c_uint64, C.c_int, C.c_float, _P, _P, _P, _P,
                                C.c_int]
        L.orc_occluded.restype = C.c_int
        L.orc_occluded.argtypes = [C.POINTER(_abi.HjkScene), _P, C.c_uint64, C.c_int, C.c_float, _P, C.c_int]
        L.orc_render.restype = C.c_int
        L.orc_render.argtypes = [C.POINTER(_abi.HjkScene), _P, C.c_uint64, C.POINTER(OrcParams), _P,
                                 C.POINTER(OrcStats), C.c_int]
        L.orc_integrate_frame.restype = C.c_int
        L.orc_integrate_frame.argtypes = [C.POINTER(_abi.HjkScene), _P, C.c_uint64, C.POINTER(OrcParams), _P,
                                          C.POINTER(OrcStats), C.c_int]
        L.orc_reconstruct_frame.restype = C.c_int
        L.orc_reconstruct_frame.argtypes = [_P, C.c_uint64, C.POINTER(OrcParams), _P, _P, _P, _P, C.c_int]
        L.orc_trace_path.restype = C.c_int
        L.orc_trace_path.argtypes = [C.POINTER(_abi.HjkScene), _P, C.c_uint32, C.c_uint32,
                                     C.POINTER(OrcParams), _P, C.c_int]
        L.orc_hardware_threads.restype = C.c_int
        L.orc_math_eval.restype = None
        L.orc_math_eval.argtypes = [C.c_int, _P, _P, _P, C.c_uint64]
        _oracle = L
    return _oracle


def hosttest() -> C.CDLL:
    global _hosttest
    if _hosttest is None:
        _make(NATIVE_DIR)
        L = C.CDLL(os.path.join(NATIVE_DIR, "libhjk_hosttest.so"))
        L.ht_create.restype = _P
        L.ht_create.argtypes = [C.POINTER(_abi.HjkScene), C.c_float, C.c_char_p, C.c_int]
        L.ht_destroy.restype = None
        L.ht_destroy.argtypes = [_P]
        L.ht_bvh_stats.restype = None
        L.ht_bvh_stats.argtypes = [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                   C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.ht_trace.restype = None
        L.ht_trace.argtypes = [_P, _P, C.c_uint64, C.c_int, C.c_float, _P, _P, _P]
        L.ht_render.restype = C.c_int
        L.ht_render.argtypes = [_P, _P, C.c_uint64, C.POINTER(_abi.HjkParams), _P, _P, _P]
        L.ht_denoise.restype = C.c_int
        L.ht_denoise.argtypes = [_P, C.c_uint64, C.POINTER(_abi.HjkParams), _P, _P, _P, _P]
        L.ht_set_exact.restype = None
        L.ht_set_exact.argtypes = [C.c_int]
        L.ht_unresolved.restype = C.c_uint64
        L.ht_trace_path.restype = C.c_int
        L.ht_trace_path.argtypes = [_P, _P, C.c_uint32, C.c_uint32, C.POINTER(_abi.HjkParams), _P, C.c_int, _P]
        L.ht_math_eval.restype = None
        L.ht_math_eval.argtypes = [C.c_int, _P, _P, _P, C.c_uint64]
        # the harness links the host front-end sources too (scene loading without nvcc)
        for name in ("hjk_host_scene_from_obj", "hjk_host_scene_terrain", "hjk_host_scene_spheres",
                     "hjk_host_scene_view", "hjk_host_scene_free", "hjk_host_last_error",
                     "hjk_host_generate_blocks", "hjk_host_write_exr"):
            res, args = _abi._SIGNATURES[name]
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _hosttest = L
    return _hosttest


def ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def math_eval(lib_fn, fn: int, a, b=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = a if b is None else np.ascontiguousarray(b, dtype=np.float32)
    out = np.empty_like(a)
    lib_fn(fn, ptr(a), ptr(b), ptr(out), C.c_uint64(a.size))
    return out


class HostScene:
    """A compiled scene produced by the host front-end (via whichever library `lib` is)."""

    def __init__(self, lib, handle):
        self.lib, self.handle = lib, handle
        self.view = _abi.HjkScene()
        rc = lib.hjk_host_scene_view(handle, C.byref(self.view))
        assert rc == 0

    @classmethod
    def from_obj(cls, lib, path=CBOX_OBJ, put_spheres=False, with_bvh2=True):
        h = _P()
        rc = lib.hjk_host_scene_from_obj(path.encode(), int(put_spheres), int(with_bvh2), C.byref(h))
        if rc != 0:
            raise RuntimeError(lib.hjk_host_last_error().decode())
        return cls(lib, h)

    @classmethod
    def terrain(cls, lib, grid_n, seed=7, with_bvh2=True):
        h = _P()
        rc = lib.hjk_host_scene_terrain(grid_n, seed, int(with_bvh2), C.byref(h))
        if rc != 0:
            raise RuntimeError(lib.hjk_host_last_error().decode())
        return cls(lib, h)

    @classmethod
    def spheres(cls, lib, lattice_n, seed=7, with_bvh2=True):
        h = _P()
        rc = lib.hjk_host_scene_spheres(lattice_n, seed, int(with_bvh2), C.byref(h))
        if rc != 0:
            raise RuntimeError(lib.hjk_host_last_error().decode())
        return cls(lib, h)

    def array(self, name: str) -> np.ndarray:
        arr: _abi.HjkArray = getattr(self.view, name)
        size, dtype, per = _abi.SCENE_ELEM[name]
        if arr.count == 0 or not arr.ptr:
            return np.zeros((0, per), dtype=dtype)
        buf = (C.c_uint8 * (arr.count * size)).from_address(arr.ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(arr.count, per)

    @property
    def info(self) -> _abi.HjkSceneInfo:
        return _abi.HjkSceneInfo.from_address(self.view.scene.ptr)

    def close(self):
        if self.handle:
            self.lib.hjk_host_scene_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def generate_blocks(lib, width, height, spp, block_size=128, root_seed=0x48494A494B49) -> np.ndarray:
    n = lib.hjk_host_generate_blocks(width, height, block_size, spp, root_seed, None, 0)
    blocks = np.zeros(n, dtype=_abi.BLOCK_DTYPE)
    lib.hjk_host_generate_blocks(width, height, block_size, spp, root_seed, ptr(blocks), n)
    return blocks


def camera_rays(scene: HostScene, width, height, offset=(0.5, 0.5), eps=1e-4) -> np.ndarray:
    """Primary rays of every pixel, from the ORACLE's camera (render.glsl:26-36)."""
    L = oracle()
    rays = np.zeros(width * height, dtype=_abi.RAY_DTYPE)
    one = np.zeros(1, dtype=_abi.RAY_DTYPE)
    k = 0
    for y in range(height):
        for x in range(width):
            L.orc_camera_ray(scene.view.scene.ptr, x + offset[0], y + offset[1], float(width), float(height),
                             eps, ptr(one))
            rays[k] = one[0]
            k += 1
    return rays


def orc_params(max_bounces=1000, rr_start=3, radius=2, stddev=0.5, eps=1e-4, use_bvh=1, block_size=128,
               skip_recon=0) -> OrcParams:
    return OrcParams(max_bounces, rr_start, radius, stddev, eps, use_bvh, block_size, skip_recon)


def hjk_params(max_bounces=1000, rr_start=3, radius=2, stddev=0.5, eps=1e-4, flags=0) -> _abi.HjkParams:
    return _abi.HjkParams(max_bounces, rr_start, radius, stddev, eps, flags)


class CustomScene:
    """A compiled scene assembled from numpy arrays in the reference's binding layouts.
    Shape order = [spheres | quads | triangles]; `materials` = one (tag, index) pair per shape."""

    def __init__(self, camera, spheres=(), quads=(), triangles=(), vertices=(), materials=(), diffuse=(),
                 diffusecb=(), dielectric=(), emissive=()):
        f32 = np.float32
        self.arrays = {
            "spheres": np.asarray(spheres, f32).reshape(-1, 4),
            "quads": np.asarray(quads, f32).reshape(-1, 12),
            "triangles": np.asarray(triangles, np.uint32).reshape(-1, 3),
            "vertices": np.asarray(vertices, f32).reshape(-1, 8),
            "diffuse": np.asarray(diffuse, f32).reshape(-1, 4),
            "diffusecb": np.asarray(diffusecb, f32).reshape(-1, 8),
            "dielectric": np.asarray(dielectric, f32).reshape(-1, 4),
            "emissive": np.asarray(emissive, f32).reshape(-1, 4),
        }
        mats = np.array([(t << 24) | i for t, i in materials], np.uint32)
        self.arrays["materials"] = mats
        em_shapes = [k for k, (t, _) in enumerate(materials) if t == _abi.MAT_EMISSIVE]
        em = np.zeros((len(em_shapes), 4), f32)
        if em_shapes:
            em.view(np.uint32)[:, 0] = em_shapes
            em[:, 1] = f32(1.0) / f32(len(em_shapes))
            em[:, 2] = np.cumsum(em[:, 1], dtype=f32)
        self.arrays["emitters"] = em
        self.info_struct = _abi.HjkSceneInfo()
        pos, rot, fov = camera
        for k in range(3):
            self.info_struct.camera.position[k] = pos[k]
        for k in range(4):
            self.info_struct.camera.rotation[k] = rot[k]
        self.info_struct.camera.fov = fov
        self.info_struct.num_spheres = len(self.arrays["spheres"])
        self.info_struct.num_quads = len(self.arrays["quads"])
        self.info_struct.num_triangles = len(self.arrays["triangles"])
        self.info_struct.num_emitters = len(em_shapes)
        self.view = _abi.HjkScene()
        self.view.scene = _abi.HjkArray(C.addressof(self.info_struct), 1)
        self.view.bvh = _abi.HjkArray(None, 0)
        for name, a in self.arrays.items():
            a = np.ascontiguousarray(a)
            self.arrays[name] = a
            setattr(self.view, name, _abi.HjkArray(a.ctypes.data if a.size else None, len(a)))

    @property
    def info(self):
        return self.info_struct

    def array(self, name):
        return self.arrays[name]


def quad_room_scene():
    """A closed room of six quads (checkerboard floor, diffuse walls) with an emissive quad under the
    ceiling, a tinted-glass sphere (non-zero extinction), a mirror sphere and one diffuse triangle:
    exercises shapes/quad.glsl, the quad emitter, DielectricMaterial::tinted and Beer-Lambert."""
    D, CB, MI, DI, EM = _abi.MAT_DIFFUSE, _abi.MAT_DIFFUSECBOARD, _abi.MAT_MIRROR, _abi.MAT_DIELECTRIC, _abi.MAT_EMISSIVE
    q = lambda o, e1, e2: [*o, 0, *e1, 0, *e2, 0]
    quads = [
        q((-1, 0, -1), (0, 0, 2), (2, 0, 0)),      # floor (normal +y)
        q((-1, 2, -1), (2, 0, 0), (0, 0, 2)),      # ceiling (normal -y)
        q((-1, 0, -1), (2, 0, 0), (0, 2, 0)),      # back wall (normal +z)
        q((-1, 0, -1), (0, 2, 0), (0, 0, 2)),      # left wall (normal +x)
        q((1, 0, -1), (0, 0, 2), (0, 2, 0)),       # right wall (normal -x)
        q((-0.3, 1.98, -0.3), (0.6, 0, 0), (0, 0, 0.6)),  # light, facing down
    ]
    spheres = [(-0.4, 0.35, -0.2, 0.35), (0.45, 0.3, 0.1, 0.3)]
    vertices = [(-0.2, 0.0, 0.6, 0, 0, 1, 0, 0), (0.3, 0.0, 0.7, 1, 0, 1, 0, 0), (0.0, 0.5, 0.5, 0, 0, 1, 0, 1)]
    triangles = [(0, 1, 2)]
    materials = [(DI, 0), (MI, 0), (CB, 0), (D, 0), (D, 0), (D, 1), (D, 2), (EM, 0), (D, 0)]
    half = np.deg2rad(-6.0) / 2
    camera = ((0.0, 1.0, 4.2), (float(np.sin(half)), 0.0, 0.0, float(np.cos(half))), 38.0)
    return CustomScene(camera, spheres=spheres, quads=quads, triangles=triangles, vertices=vertices,
                       materials=materials, diffuse=[(0.7, 0.7, 0.7, 0), (0.7, 0.2, 0.2, 0), (0.2, 0.6, 0.25, 0)],
                       diffusecb=[(0.8, 0.8, 0.8, 0.25, 0.15, 0.15, 0.3, 0.25)],
                       dielectric=[(0.9, 0.3, 0.2, 1.5)], emissive=[(12, 12, 12, 0)])


def random_scene(seed: int, n_tris: int = 60, n_spheres: int = 3, n_quads: int = 2):
    """A random soup in the binding layouts: triangles with random normals/uvs, spheres, quads, every
    material tag (tinted glass included), at least one emitter of each emissive-capable shape kind."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    D, CB, MI, DI, EM = _abi.MAT_DIFFUSE, _abi.MAT_DIFFUSECBOARD, _abi.MAT_MIRROR, _abi.MAT_DIELECTRIC, _abi.MAT_EMISSIVE
    verts = np.zeros((3 * n_tris, 8), f32)
    centres = rng.uniform(-1.5, 1.5, (n_tris, 1, 3))
    verts[:, :3] = (centres + rng.normal(0, 0.5, (n_tris, 3, 3))).reshape(-1, 3)
    nrm = rng.normal(0, 1, (3 * n_tris, 3))
    verts[:, 4:7] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    verts[:, 3] = rng.uniform(0, 4, 3 * n_tris)
    verts[:, 7] = rng.uniform(0, 4, 3 * n_tris)
    tris = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
    spheres = np.concatenate([rng.uniform(-1.5, 1.5, (n_spheres, 3)), rng.uniform(0.2, 0.5, (n_spheres, 1))], 1)
    quads = []
    for _ in range(n_quads):
        o = rng.uniform(-2, 2, 3)
        e1 = rng.normal(0, 1, 3)
        e2 = np.cross(e1, rng.normal(0, 1, 3))
        quads.append([*o, 0, *e1, 0, *(e2 / np.linalg.norm(e2) * rng.uniform(0.5, 1.5)), 0])
    n_shapes = n_spheres + n_quads + n_tris
    diffuse = np.concatenate([rng.uniform(0.2, 0.9, (4, 3)), np.zeros((4, 1))], 1)
    diffusecb = np.array([[0.8, 0.8, 0.7, 0.3, 0.2, 0.3, 0.2, 0.4], [0.6, 0.1, 0.1, 0.15, 0.9, 0.9, 0.9, 0.2]])
    dielectric = np.array([[0, 0, 0, 1.5], [0.6, 0.2, 0.9, 1.33]])
    emissive = np.array([[8, 8, 8, 0], [4, 6, 9, 0]])
    choices = [(D, 0), (D, 1), (D, 2), (D, 3), (CB, 0), (CB, 1), (MI, 0), (DI, 0), (DI, 1), (EM, 0), (EM, 1)]
    weights = np.array([3, 3, 3, 3, 2, 2, 2, 2, 2, 1, 1], float)
    mats = [choices[k] for k in rng.choice(len(choices), n_shapes, p=weights / weights.sum())]
    mats[0] = (EM, 0)                       # an emissive sphere
    mats[n_spheres] = (EM, 1)               # an emissive quad
    mats[n_spheres + n_quads] = (EM, 0)     # an emissive triangle
    half = rng.uniform(-0.3, 0.3)
    camera = ((0.0, 0.0, 6.0), (float(np.sin(half)), 0.0, 0.0, float(np.cos(half))), 45.0)
    return CustomScene(camera, spheres=spheres, quads=quads, triangles=tris, vertices=verts, materials=mats,
                       diffuse=diffuse, diffusecb=diffusecb, dielectric=dielectric, emissive=emissive)
