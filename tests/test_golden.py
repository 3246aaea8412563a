"""Committed golden vectors (tests/golden/, generated from the oracle by make_golden.py):
the oracle must keep reproducing them (pins the checker against drift), the product's kernel logic
compiled for the host must match them (CPU), and the CUDA path must match them (GPU)."""
import ctypes as C
import os

import numpy as np
import pytest

import _libs
from hijiki_b200 import _abi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIRST_HIT = [("cbox_first_hit_64x48.npz", "cbox"), ("cbox_spheres_first_hit_64x48.npz", "cbox_spheres")]
RENDER = [("cbox_acc_72x48_2spp_b8.npz", "cbox"), ("cbox_spheres_acc_72x48_2spp_b1000.npz", "cbox_spheres"),
          ("lattice3_acc_64x40_2spp_b16.npz", "lattice3")]


def _scene(hosttest, request, name):
    if name == "lattice3":
        return _libs.HostScene.spheres(hosttest, 3)
    return request.getfixturevalue(name)


def _load(name):
    return np.load(os.path.join(GOLD, name))


@pytest.mark.parametrize("fname,scene_name", FIRST_HIT)
def test_oracle_and_kernel_logic_reproduce_first_hit_golden(oracle, hosttest, request, fname, scene_name):
    g = _load(fname)
    scene = _scene(hosttest, request, scene_name)
    rays = np.ascontiguousarray(g["rays"]).view(_abi.RAY_DTYPE).reshape(-1)
    n = rays.size
    ids, t = np.zeros(n, np.int32), np.zeros(n, np.float32)
    assert oracle.orc_trace(C.byref(scene.view), _libs.ptr(rays), n, 0, 1e-4, _libs.ptr(ids), _libs.ptr(t), None,
                            None, 0) == 0
    assert np.array_equal(ids, g["ids"]) and np.array_equal(t.view(np.uint32), g["t"].view(np.uint32))
    err = C.create_string_buffer(256)
    h = hosttest.ht_create(C.byref(scene.view), 1e-5, err, 256)
    hosttest.ht_trace(h, _libs.ptr(rays), n, 0, 1e-4, _libs.ptr(ids), _libs.ptr(t), None)
    hosttest.ht_destroy(h)
    keep = g["tie"] == 0
    assert np.array_equal(ids[keep], g["ids"][keep])
    hit = keep & (g["ids"] >= 0)
    assert np.array_equal(t[hit].view(np.uint32), g["t"][hit].view(np.uint32))


@pytest.mark.parametrize("fname,scene_name", RENDER)
def test_oracle_and_kernel_logic_reproduce_render_golden(oracle, hosttest, request, fname, scene_name):
    g = _load(fname)
    scene = _scene(hosttest, request, scene_name)
    blocks = np.ascontiguousarray(g["blocks"]).view(_abi.BLOCK_DTYPE).reshape(-1)
    max_bounces, bs, mode = (int(v) for v in g["params"])
    acc = np.zeros_like(g["acc"])
    st = _libs.OrcStats()
    op = _libs.orc_params(max_bounces=max_bounces, use_bvh=mode, block_size=bs)
    assert oracle.orc_render(C.byref(scene.view), _libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(acc),
                             C.byref(st), 0) == 0
    assert np.array_equal(acc.view(np.uint32), g["acc"].view(np.uint32))
    assert [st.n_paths, st.n_extension_rays, st.n_shadow_rays] == g["counts"].tolist()
    err = C.create_string_buffer(256)
    h = hosttest.ht_create(C.byref(scene.view), 1e-5, err, 256)
    acc_h = np.zeros_like(g["acc"])
    hp = _libs.hjk_params(max_bounces=max_bounces)
    assert hosttest.ht_render(h, _libs.ptr(blocks), blocks.size, C.byref(hp), _libs.ptr(acc_h), None, None) == 0
    hosttest.ht_destroy(h)
    diff = (acc_h.view(np.uint32) != g["acc"].view(np.uint32)).any(axis=2)
    assert diff.sum() <= 25  # at most one tie-affected sample


@pytest.mark.gpu
@pytest.mark.parametrize("fname,scene_name", FIRST_HIT)
def test_cuda_first_hit_matches_golden(gpu_ctx, fname, scene_name):
    import hijiki_b200 as hj
    g = _load(fname)
    gpu_ctx.scene_upload(hj.Scene.from_obj(_libs.CBOX_OBJ, put_cbox_spheres=scene_name == "cbox_spheres").compile())
    rays = np.ascontiguousarray(g["rays"]).view(_abi.RAY_DTYPE).reshape(-1)
    ids, t, uv = gpu_ctx.trace_first_hit(rays)
    keep = g["tie"] == 0
    assert np.array_equal(ids[keep], g["ids"][keep])
    hit = keep & (g["ids"] >= 0)
    assert np.array_equal(t[hit].view(np.uint32), g["t"][hit].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("fname,scene_name", RENDER)
def test_cuda_render_matches_golden(gpu_ctx, fname, scene_name):
    import hijiki_b200 as hj
    g = _load(fname)
    scene = hj.Scene.spheres(3) if scene_name == "lattice3" else \
        hj.Scene.from_obj(_libs.CBOX_OBJ, put_cbox_spheres=scene_name == "cbox_spheres")
    gpu_ctx.scene_upload(scene.compile())
    blocks = np.ascontiguousarray(g["blocks"]).view(_abi.BLOCK_DTYPE).reshape(-1)
    max_bounces = int(g["params"][0])
    h, w = g["acc"].shape[:2]
    gpu_ctx.frame_begin(w, h)
    st = gpu_ctx.render(blocks, hj.make_params(max_bounces=max_bounces))
    acc = gpu_ctx.readback(normalise=False)
    diff = (acc.view(np.uint32) != g["acc"].view(np.uint32)).any(axis=2)
    assert diff.sum() <= 25
    assert st.n_paths == int(g["counts"][0])
    assert abs(st.n_extension_rays - int(g["counts"][1])) <= 20
