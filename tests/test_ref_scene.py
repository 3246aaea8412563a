"""The product's scene front-end (obj_loader.cpp + scene_compile.cpp behind hjk_host_scene_from_obj) against the
independent Python restatement of Scene::from_obj / Scene::compile in tests/ref_scene.py, byte for byte, plus
hand-checked facts of scenes/cbox that do not come from either loader."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import _libs
import ref_scene
from hijiki_b200 import _abi

ARRAYS = ("spheres", "quads", "triangles", "vertices", "materials", "emitters", "diffuse", "diffusecb", "dielectric",
          "emissive")


@pytest.mark.parametrize("put_spheres", [False, True])
def test_product_loader_matches_the_independent_restatement(hosttest, put_spheres):
    ours = _libs.HostScene.from_obj(hosttest, _libs.CBOX_OBJ, put_spheres=put_spheres, with_bvh2=False)
    ref = ref_scene.RefScene(_libs.CBOX_OBJ, put_cbox_spheres=put_spheres)
    assert bytes(ours.info) == bytes(ref.info)  # camera + the four counts (SceneBufferInfo, src/main.rs:400-408)
    for name in ARRAYS:
        a, b = ours.array(name), ref.array(name)
        assert a.shape[0] == b.shape[0], name
        assert a.tobytes() == b.reshape(a.shape).tobytes(), name


def test_cbox_facts_checked_by_hand():
    """Read off scenes/cbox/cbox.obj and cbox.mtl with grep/awk, not with a loader: 6320 triangles + 6 quads in
    the file = 6332 triangles; materials in MTL order floor, light, porcelain, wall_blue, wall_gray, wall_red;
    the one `light*` material becomes the only emissive entry with Ke = 15; the light model is one quad = two
    triangles, so two emitters with pdf 1/2 and cdf 1/2, 1."""
    ref = ref_scene.RefScene(_libs.CBOX_OBJ)
    assert ref.info.num_triangles == 6320 + 2 * 6
    assert (ref.info.num_spheres, ref.info.num_quads, ref.info.num_emitters) == (0, 0, 2)
    assert ref.array("diffuse").shape[0] == 5 and ref.array("emissive").shape[0] == 1
    assert ref.array("emissive")[0].tolist() == [15.0, 15.0, 15.0, 0.0]
    assert np.allclose(ref.array("diffuse")[0, :3], [0.455928, 0.446495, 0.427629])  # floor, first in the MTL
    em = ref.array("emitters")
    assert em[:, 1].tolist() == [0.5, 0.5] and em[:, 2].tolist() == [0.5, 1.0]
    shapes = em.view(np.uint32)[:, 0]
    assert (ref.array("materials")[shapes] >> 24 == 4).all() and shapes[1] == shapes[0] + 1
    # models appear in file order: the teapot's 6320 triangles first (porcelain = diffuse index 1), then the walls
    mats = ref.array("materials")
    assert (mats[:6320] == (0 << 24) + 1).all()
    # the emitter triangles belong to the 4th model (`o light`): rightWall, leftWall (2 triangles each) precede it
    assert shapes[0] == 6320 + 4
    # pinned digest of the arrays the reference would hand its shaders (all but `bvh`)
    h = hashlib.sha256()
    for name in ARRAYS:
        h.update(ref.array(name).tobytes())
    assert h.hexdigest() == CBOX_SHA256


CBOX_SHA256 = "4bd955f272903d425fd02a8ab1fc1d9ff6be7e981d039fffc7ae48abc07806fa"


def test_oracle_renders_the_same_frame_from_either_loader(oracle, hosttest):
    """The oracle fed by the product's loader and by the independent one: same frame, bit for bit."""
    ours = _libs.HostScene.from_obj(hosttest, _libs.CBOX_OBJ, put_spheres=True, with_bvh2=False)
    ref = ref_scene.RefScene(_libs.CBOX_OBJ, put_cbox_spheres=True)
    w, h = 48, 32
    blocks = _libs.generate_blocks(hosttest, w, h, 1, 64)
    out = []
    for scene in (ours, ref):
        acc = np.zeros((h, w, 4), np.float32)
        op = _libs.orc_params(max_bounces=6, use_bvh=3, block_size=64)
        assert oracle.orc_render(C.byref(scene.view), _libs.ptr(blocks), blocks.size, C.byref(op), _libs.ptr(acc), None,
                                 0) == 0
        out.append(acc)
    assert np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32))
