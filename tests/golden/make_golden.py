"""Generates the committed golden vectors from the ORACLE (the reference itself cannot run in this
image and ships no fixtures).  Re-run: python tests/golden/make_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _libs  # noqa: E402

O, H = _libs.oracle(), _libs.hosttest()


def first_hit(scene, w, h, mode):
    rays = _libs.camera_rays(scene, w, h)
    n = rays.size
    ids, t, uv, tie = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros((n, 2), np.float32), np.zeros(n, np.uint8)
    assert O.orc_trace(C.byref(scene.view), _libs.ptr(rays), n, mode, 1e-4, _libs.ptr(ids), _libs.ptr(t), _libs.ptr(uv),
                       _libs.ptr(tie), 0) == 0
    return dict(rays=rays.view(np.float32).reshape(n, 8), ids=ids, t=t, uv=uv, tie=tie)


def render(scene, w, h, spp, bs, max_bounces, mode):
    blocks = _libs.generate_blocks(H, w, h, spp, block_size=bs)
    acc = np.zeros((h, w, 4), np.float32)
    st = _libs.OrcStats()
    op = _libs.orc_params(max_bounces=max_bounces, use_bvh=mode, block_size=bs)
    assert O.orc_render(C.byref(scene.view), _libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(acc), C.byref(st), 0) == 0
    return dict(blocks=blocks.view(np.uint8).reshape(blocks.size, 40), acc=acc,
                counts=np.array([st.n_paths, st.n_extension_rays, st.n_shadow_rays], np.uint64),
                params=np.array([max_bounces, bs, mode], np.int64))


cbox = _libs.HostScene.from_obj(H, _libs.CBOX_OBJ, False, True)
cbox_s = _libs.HostScene.from_obj(H, _libs.CBOX_OBJ, True, True)
lattice = _libs.HostScene.spheres(H, 3)
np.savez_compressed(os.path.join(HERE, "cbox_first_hit_64x48.npz"), **first_hit(cbox, 64, 48, 0))
np.savez_compressed(os.path.join(HERE, "cbox_spheres_first_hit_64x48.npz"), **first_hit(cbox_s, 64, 48, 0))
np.savez_compressed(os.path.join(HERE, "cbox_acc_72x48_2spp_b8.npz"), **render(cbox, 72, 48, 2, 64, 8, 0))
np.savez_compressed(os.path.join(HERE, "cbox_spheres_acc_72x48_2spp_b1000.npz"), **render(cbox_s, 72, 48, 2, 64, 1000, 0))
np.savez_compressed(os.path.join(HERE, "lattice3_acc_64x40_2spp_b16.npz"), **render(lattice, 64, 40, 2, 64, 16, 2))
print("golden vectors written to", HERE)
