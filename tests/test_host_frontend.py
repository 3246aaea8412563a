"""Host logic (no GPU): block generator, pass planning, EXR writer, sample-pass split, and that the
C-ABI library loads and exports every symbol include/hijiki_b200.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _libs
import hijiki_b200 as hj
from hijiki_b200 import _abi


def test_library_exports_every_declared_symbol():
    lib = _abi.load()  # raises if the in-tree .so is missing
    header = open(os.path.join(_libs.ROOT, "include", "hijiki_b200.h")).read()
    declared = set(re.findall(r"HJK_API\s+[\w\s\*]+?\b(hjk_\w+)\s*\(", header))
    assert declared, "no HJK_API declarations found"
    assert declared == set(_abi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.hjk_version()


def test_no_gpu_fails_loudly_not_silently():
    """Without a CUDA device the product refuses to run — there is no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hj.HijikiError):
        hj.Context(0)


def test_block_generator_matches_reference_iteration():
    """ImageBlockGenerator::next (src/main.rs:648-682): row-major tiles, spp passes, clipped edges,
    one sample_offset per pass — re-drawn BEFORE the last block of the pass is emitted."""
    gen = hj.ImageBlockGenerator(1920, 1080, 128, 3)
    b = gen.blocks()
    assert gen.blocks_per_pass == 135 and b.size == 405
    assert (b["id"] == np.arange(405)).all()
    first = b[:135]
    assert (first["origin"][:, 0] == np.tile(np.arange(15) * 128, 9)).all()
    assert (first["origin"][:, 1] == np.repeat(np.arange(9) * 128, 15)).all()
    assert (first["dimension"][:, 0] == 128).all()
    assert (first["dimension"][-15:, 1] == 56).all() and (first["dimension"][:-15, 1] == 128).all()
    assert (b["original_dimension"] == (1920, 1080)).all()
    so = b["sample_offset"]
    assert (so >= 0).all() and (so < 1).all()
    assert (so[:134] == so[0]).all()
    assert (so[134] == so[135]).all() and not (so[134] == so[0]).all()  # the quirk
    assert len(set(b["seed"].tolist())) > 400
    # recorded stream: same root seed, same list
    assert (hj.ImageBlockGenerator(1920, 1080, 128, 3).blocks() == b).all()
    assert not (hj.ImageBlockGenerator(1920, 1080, 128, 3, root_seed=1).blocks()["seed"] == b["seed"]).all()
    with pytest.raises(ValueError):
        hj.ImageBlockGenerator(100, 100, 100, 1)


def test_split_passes_partitions_every_pass_once():
    gen = hj.ImageBlockGenerator(300, 200, 64, 7)
    b = gen.blocks()
    parts = [hj.split_passes(b, gen.blocks_per_pass, r, 4) for r in range(4)]
    ids = np.concatenate([p["id"] for p in parts])
    assert sorted(ids.tolist()) == list(range(b.size))
    assert [p.size // gen.blocks_per_pass for p in parts] == [2, 2, 2, 1]
    assert (hj.split_passes(b, gen.blocks_per_pass, 0, 1) == b).all()


def test_exr_writer_roundtrip(tmp_path):
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    cv2 = pytest.importorskip("cv2")
    lib = _abi.load()
    rng = np.random.default_rng(3)
    img = rng.random((37, 53, 4), dtype=np.float32)
    path = str(tmp_path / "out.exr")
    assert lib.hjk_host_write_exr(path.encode(), _abi.as_ptr(img), 53, 37, img.strides[0]) == 0
    back = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if back is None:
        pytest.skip("this cv2 build cannot read EXR")
    assert back.shape == (37, 53, 3)
    assert np.array_equal(back[..., ::-1], img[..., :3])  # cv2 returns BGR
    assert lib.hjk_host_write_exr(b"/nonexistent_dir/x.exr", _abi.as_ptr(img), 53, 37, img.strides[0]) == -7


def test_scene_errors_are_reported_not_thrown():
    lib = _abi.load()
    h = C.c_void_p()
    assert lib.hjk_host_scene_from_obj(b"/no/such/file.obj", 0, 0, C.byref(h)) != 0
    assert lib.hjk_host_last_error()
    assert lib.hjk_host_scene_terrain(0, 1, 0, C.byref(h)) == -1
    with pytest.raises(hj.HijikiError):
        hj.Scene.from_obj("/no/such/file.obj").compile()


@pytest.mark.parametrize("kind", ["cbox", "cbox_spheres", "spheres", "terrain"])
def test_wide_bvh_is_structurally_valid(kind):
    scene = {"cbox": hj.Scene.from_obj(_libs.CBOX_OBJ), "cbox_spheres": hj.Scene.from_obj(_libs.CBOX_OBJ, True),
             "spheres": hj.Scene.spheres(8), "terrain": hj.Scene.terrain(96)}[kind].compile()
    st = scene.bvh_stats()
    info = scene.info
    n = info.num_spheres + info.num_quads + info.num_triangles
    assert st["valid"], _abi.load().hjk_host_last_error()
    assert st["prims"] == n
    assert st["depth"] <= 32 and st["nodes"] < n
    assert st["bytes"] == st["nodes"] * 80 + st["prims"] * 48


def test_wide_bvh_does_not_depend_on_the_thread_count():
    """The host builder runs its SAH binning, collapse sweep and emission on every host thread; nodes and
    primitive records must come out bit-identical on 1, 3 and all threads (1.16 M triangles: large enough for
    the all-threads binning of nodes over 2^20 primitives, the per-subtree sweeps and the emission arenas)."""
    for scene in (hj.Scene.from_obj(_libs.CBOX_OBJ).compile(), hj.Scene.terrain(760).compile()):
        digests = {scene.bvh_digest(n) for n in (1, 3, 0)}
        assert len(digests) == 1, digests


def test_synthetic_scene_shapes():
    sp = hj.Scene.spheres(8).compile()
    assert sp.info.num_spheres == 512 and sp.info.num_triangles == 4 and sp.info.num_emitters == 2
    tags = sp.array("materials")[:512, 0] >> 24
    assert set(tags.tolist()) == {_abi.MAT_MIRROR, _abi.MAT_DIELECTRIC}
    assert np.allclose(sp.array("dielectric")[0], (0, 0, 0, 1.5))
    r = sp.array("spheres")[:, 3]
    assert r.min() >= 0.25 and r.max() <= 0.42
    te = hj.Scene.terrain(64).compile()
    assert te.info.num_triangles == 2 * 64 * 64 + 32 and te.info.num_emitters == 32
    assert te.array("vertices").shape[0] == 65 * 65 + 16 * 4


def _build_c_example(tmp_path):
    import subprocess
    exe = str(tmp_path / "render_c")
    lib_dir = os.path.dirname(_abi.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(_libs.ROOT, "include"),
           os.path.join(_libs.ROOT, "examples", "render_c.c"), "-L", lib_dir, "-lhijiki_b200", f"-Wl,-rpath,{lib_dir}",
           "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _build_c_example_static(tmp_path):
    """The link a Rust host gets: libhijiki_b200.a + exactly the libraries bindings/rust/build.rs prints as
    cargo:rustc-link-lib / rustc-link-search (parsed from build.rs, so the two cannot drift apart)."""
    import re
    import subprocess
    src = open(os.path.join(_libs.ROOT, "bindings", "rust", "build.rs")).read()
    search = re.findall(r'rustc-link-search=native=([^"{}]+)"', src)      # literal paths (the OUT_DIR one is ours)
    static = re.findall(r'rustc-link-lib=static=(\w+)', src)
    dynamic = re.search(r'for l in \[([^\]]+)\]', src).group(1).replace('"', "").replace(" ", "").split(",")
    assert static[0] == "hijiki_b200" and "cudart_static" in static and "stdc++" in dynamic
    exe = str(tmp_path / "render_c_static")
    lib_dir = os.path.dirname(_abi.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-I", os.path.join(_libs.ROOT, "include"), os.path.join(_libs.ROOT, "examples", "render_c.c"),
           "-L", lib_dir] + [f"-L{p}" for p in search] + [f"-l:lib{n}.a" for n in static] + [f"-l{n}" for n in dynamic] + \
          ["-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "hijiki" not in ldd and "cudart" not in ldd  # both are inside the executable
    return exe


def test_c_host_links_the_static_library_with_the_rust_link_line(tmp_path):
    import subprocess
    exe = _build_c_example_static(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run itself is covered by the gpu-marked test")
    r = subprocess.run([exe, _libs.CBOX_OBJ, "64", "48", "1", str(tmp_path / "o.exr")], capture_output=True, text=True)
    assert r.returncode == 1 and "hjk_create" in r.stderr


@pytest.mark.gpu
def test_statically_linked_c_host_renders(tmp_path):
    import subprocess
    exe = _build_c_example_static(tmp_path)
    out = tmp_path / "o.exr"
    r = subprocess.run([exe, _libs.CBOX_OBJ, "128", "96", "2", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Integrated 24576 paths" in r.stdout and out.stat().st_size > 128 * 96 * 12


def test_c_host_compiles_against_the_header_and_fails_loudly_without_a_gpu(tmp_path):
    """include/hijiki_b200.h is plain C99 (-pedantic -Werror) and examples/render_c.c — the reference's main()
    over the C ABI — links against the library; without a CUDA device it stops at hjk_create with a message."""
    import subprocess
    exe = _build_c_example(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run itself is covered by the gpu-marked test")
    r = subprocess.run([exe, _libs.CBOX_OBJ, "64", "48", "1", str(tmp_path / "o.exr")], capture_output=True, text=True)
    assert r.returncode == 1 and "hjk_create" in r.stderr


@pytest.mark.gpu
def test_c_host_renders(tmp_path):
    import subprocess
    exe = _build_c_example(tmp_path)
    out = tmp_path / "o.exr"
    r = subprocess.run([exe, _libs.CBOX_OBJ, "128", "96", "2", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Integrated 24576 paths" in r.stdout and out.stat().st_size > 128 * 96 * 12
