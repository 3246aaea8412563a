"""ctypes mirror of ``include/hijiki_b200.h`` — the C ABI that replaces the wgpu render loop of
the reference (reference src/main.rs:760-898, 907-1004, 1167-1423).

The shared library is built in-tree by ``__graft_entry__.build()`` (or ``make -C
hijiki_b200/csrc``).  There is no fallback: if it is missing, importing the product API
raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhijiki_b200.so")


# ---------------------------------------------------------------- POD layouts
class HjkCamera(C.Structure):
    _fields_ = [("position", C.c_float * 4), ("rotation", C.c_float * 4), ("fov", C.c_float),
                ("_pad", C.c_float * 3)]


class HjkSceneInfo(C.Structure):
    _fields_ = [("camera", HjkCamera), ("num_spheres", C.c_uint32), ("num_quads", C.c_uint32),
                ("num_triangles", C.c_uint32), ("num_emitters", C.c_uint32)]


class HjkArray(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("count", C.c_uint64)]


SCENE_FIELDS = ("scene", "bvh", "spheres", "quads", "triangles", "vertices", "materials", "emitters",
                "diffuse", "diffusecb", "dielectric", "emissive")
# element size in bytes / numpy view dtype+shape of each binding
SCENE_ELEM = {
    "scene": (64, np.uint8, 64), "bvh": (32, np.float32, 8), "spheres": (16, np.float32, 4),
    "quads": (48, np.float32, 12), "triangles": (12, np.uint32, 3), "vertices": (32, np.float32, 8),
    "materials": (4, np.uint32, 1), "emitters": (16, np.float32, 4), "diffuse": (16, np.float32, 4),
    "diffusecb": (32, np.float32, 8), "dielectric": (16, np.float32, 4), "emissive": (16, np.float32, 4),
}


class HjkScene(C.Structure):
    _fields_ = [(name, HjkArray) for name in SCENE_FIELDS]


class HjkImageBlock(C.Structure):
    _fields_ = [("id", C.c_uint32), ("seed", C.c_uint32), ("origin", C.c_uint32 * 2),
                ("dimension", C.c_uint32 * 2), ("original_dimension", C.c_uint32 * 2),
                ("sample_offset", C.c_float * 2)]


BLOCK_DTYPE = np.dtype([("id", "<u4"), ("seed", "<u4"), ("origin", "<u4", 2), ("dimension", "<u4", 2),
                        ("original_dimension", "<u4", 2), ("sample_offset", "<f4", 2)])
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("t_min", "<f4"), ("direction", "<f4", 3), ("t_max", "<f4")])
assert BLOCK_DTYPE.itemsize == 40 and C.sizeof(HjkImageBlock) == 40
assert RAY_DTYPE.itemsize == 32 and C.sizeof(HjkSceneInfo) == 64


class HjkParams(C.Structure):
    _fields_ = [("max_bounces", C.c_uint32), ("rr_start", C.c_uint32), ("recon_radius", C.c_uint32),
                ("recon_stddev", C.c_float), ("eps", C.c_float), ("flags", C.c_uint32)]


HJK_N_KERNEL_SLOTS = 8
KERNEL_SLOTS = ("raygen", "extend", "shade", "shadow", "recon", "other", "sort", "_7")


class HjkStats(C.Structure):
    _fields_ = [("n_paths", C.c_uint64), ("n_extension_rays", C.c_uint64), ("n_shadow_rays", C.c_uint64),
                ("ms_total", C.c_float), ("kernel_ms", C.c_float * HJK_N_KERNEL_SLOTS),
                ("n_launches", C.c_uint64)]


HJK_RENDER_ASYNC = 1
HJK_RENDER_NO_RECON = 2
HJK_RENDER_KEEP_FEATURES = 4
HJK_RENDER_EXACT_TIES = 8

MAT_DIFFUSE, MAT_DIFFUSECBOARD, MAT_MIRROR, MAT_DIELECTRIC, MAT_EMISSIVE = range(5)
MATERIAL_TAG_SHIFT = 24

STATUS = {0: "HJK_OK", -1: "HJK_ERR_INVALID_ARGUMENT", -2: "HJK_ERR_CUDA", -3: "HJK_ERR_NO_SCENE",
          -4: "HJK_ERR_NO_FRAME", -5: "HJK_ERR_UNSUPPORTED", -6: "HJK_ERR_NCCL", -7: "HJK_ERR_IO",
          -8: "HJK_ERR_OUT_OF_MEMORY"}

# every symbol include/hijiki_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_SIGNATURES = {
    "hjk_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "hjk_destroy": (C.c_int, [_P]),
    "hjk_last_error": (C.c_char_p, [_P]),
    "hjk_scene_upload": (C.c_int, [_P, C.POINTER(HjkScene)]),
    "hjk_frame_begin": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "hjk_render": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(HjkParams), C.POINTER(HjkStats)]),
    "hjk_blocks_upload": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "hjk_render_resident": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(HjkParams),
                                      C.POINTER(HjkStats)]),
    "hjk_blocks_free": (C.c_int, [_P, C.c_uint64]),
    "hjk_readback": (C.c_int, [_P, _P, C.c_uint64, C.c_int]),
    "hjk_readback_root": (C.c_int, [_P, C.c_int, _P, C.c_uint64, C.c_int]),
    "hjk_readback_begin": (C.c_int, [_P, C.c_int, _P, C.c_uint64, C.c_int]),
    "hjk_readback_wait": (C.c_int, [_P]),
    "hjk_read_features": (C.c_int, [_P, C.c_int, _P, C.c_uint64]),
    "hjk_read_intermediate": (C.c_int, [_P, C.c_int, _P]),
    "hjk_trace_first_hit": (C.c_int, [_P, _P, C.c_uint64, C.c_int, _P, _P, _P]),
    "hjk_trace_first_hit_eps": (C.c_int, [_P, _P, C.c_uint64, C.c_int, C.c_float, _P, _P, _P]),
    "hjk_denoise_pass": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint64, C.POINTER(HjkParams)]),
    "hjk_denoise_upload": (C.c_int, [_P, _P, _P, _P, C.c_uint64]),
    "hjk_denoise_resident": (C.c_int, [_P, C.POINTER(HjkParams), C.c_uint32, C.POINTER(C.c_float)]),
    "hjk_comm_unique_id": (C.c_int, [_P]),
    "hjk_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "hjk_reduce_frame": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float)]),
    "hjk_allreduce_accumulator": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "hjk_accumulator_device_ptr": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "hjk_synchronize": (C.c_int, [_P]),
    "hjk_set_stream": (C.c_int, [_P, _P]),
    "hjk_set_profiling": (C.c_int, [_P, C.c_int]),
    "hjk_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "hjk_get_info": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int64)]),
    "hjk_version": (C.c_char_p, []),
    "hjk_host_scene_from_obj": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(_P)]),
    "hjk_host_scene_terrain": (C.c_int, [C.c_uint32, C.c_uint64, C.c_int, C.POINTER(_P)]),
    "hjk_host_scene_spheres": (C.c_int, [C.c_uint32, C.c_uint64, C.c_int, C.POINTER(_P)]),
    "hjk_host_scene_view": (C.c_int, [_P, C.POINTER(HjkScene)]),
    "hjk_host_scene_free": (C.c_int, [_P]),
    "hjk_host_last_error": (C.c_char_p, []),
    "hjk_host_generate_blocks": (C.c_uint64, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                              _P, C.c_uint64]),
    "hjk_host_write_exr": (C.c_int, [C.c_char_p, _P, C.c_uint32, C.c_uint32, C.c_uint64]),
    "hjk_host_bvh_stats": (C.c_int, [C.POINTER(HjkScene), C.c_float, _P]),
    "hjk_host_bvh_digest": (C.c_int, [C.POINTER(HjkScene), C.c_float, C.c_int, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def bind(lib: C.CDLL) -> C.CDLL:
    """Attach restype/argtypes for every declared symbol (raises AttributeError if one is missing)."""
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load(path: str | None = None) -> C.CDLL:
    """Load libhijiki_b200.so.  Fails loudly: the product has no CPU or library fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("HIJIKI_B200_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise RuntimeError(
            f"hijiki_b200: CUDA library not built ({p}); run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` or `make -C hijiki_b200/csrc` — there is no fallback path")
    lib = bind(C.CDLL(p))
    if path is None:
        _lib = lib
    return lib


class HijikiError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


def as_ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)
