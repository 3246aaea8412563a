"""Two GPUs through the LIBRARY's own multi-GPU paths (no torch anywhere in them):

* two processes joined by hjk_comm_unique_id / hjk_comm_init (NCCL loaded with dlopen): rank 0 builds the wide BVH
  and broadcasts it, each rank renders the passes p = rank (mod 2), the readbacks sum the frames — to every rank
  (hjk_readback) or to rank 0 only (hjk_readback_root) — and the averaged feature buffers travel in the same
  collective;
* one process driving both devices as a group (hjk_create with two device ids, SURVEY 8b's single-process form).

The result equals the single-GPU frame up to fp32 summation order.  Skipped unless two CUDA devices are visible."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
W, H, BS, SPP, BOUNCES = 160, 96, 64, 4, 8


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, conn, out_path, spp):
    import faulthandler
    import traceback
    faulthandler.dump_traceback_later(120, exit=True)  # a hung collective must not outlive the test
    try:
        _worker_body(rank, world, conn, out_path, spp)
    except BaseException:
        with open(out_path + ".err", "w") as f:
            f.write(traceback.format_exc())
        raise


def _worker_body(rank, world, conn, out_path, spp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import hijiki_b200 as hj
    import _libs
    if rank == 0:
        uid = hj.comm_unique_id()
        conn.send(uid)
    else:
        uid = conn.recv()
    scene = hj.Scene.from_obj(_libs.CBOX_OBJ, put_cbox_spheres=True)
    gen = hj.ImageBlockGenerator(W, H, BS, spp)
    r = hj.Renderer.new(scene, gen, 128, False, device=rank, max_bounces=BOUNCES, rank=rank, world=world, comm_id=uid)
    assert r.ctx.get_info("n_ranks") == world and r.ctx.get_info("rank") == rank
    assert r.ctx.get_info("bvh_nodes") > 0  # rank 1 received the tree rank 0 built
    r.ctx.set_option("feature_buffers", 1)
    r.ctx.frame_begin(W, H)
    r.render()
    acc = r.ctx.readback(normalise=False)        # sum to every rank
    again = r.ctx.readback(normalise=False)      # a second readback does not sum a second time
    assert np.array_equal(acc, again)
    root_only = r.ctx.readback_root(0, normalise=False)
    feat = r.ctx.read_features(0)
    assert (root_only is None) == (rank != 0) and (feat is None) == (rank != 0)
    if rank == 0:
        assert np.array_equal(root_only, acc)
        np.save(out_path + ".feat.npy", feat)
    # more passes into the same frame after a reduction: the local frame was left untouched by it
    r.render()
    twice = r.ctx.readback(normalise=False)
    assert np.allclose(twice, 2 * acc, rtol=1e-6)
    np.save(out_path, acc)


def _single(spp, features=False):
    sys.path.insert(0, ROOT)
    import hijiki_b200 as hj
    import _libs
    single = hj.Renderer.new(hj.Scene.from_obj(_libs.CBOX_OBJ, put_cbox_spheres=True),
                             hj.ImageBlockGenerator(W, H, BS, spp), 128, False, max_bounces=BOUNCES)
    if features:
        single.ctx.set_option("feature_buffers", 1)
        single.ctx.frame_begin(W, H)
    single.render()
    return single.ctx.readback(normalise=False), (single.ctx.read_features() if features else None)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("spp", [SPP, 1])  # 1 spp on 2 ranks: rank 1 has no pass and still joins the collectives
def test_library_communicator_reduce(tmp_path, spp):
    ctx = mp.get_context("spawn")
    a, b = ctx.Pipe()
    outs = [str(tmp_path / f"acc{r}.npy") for r in range(2)]
    procs = [ctx.Process(target=_worker, args=(r, 2, a if r == 0 else b, outs[r], spp), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    try:
        for p in procs:
            p.join(timeout=200)
        errors = [open(o + ".err").read() for o in outs if os.path.exists(o + ".err")]
        assert not errors, "\n".join(errors)
        assert [p.exitcode for p in procs] == [0, 0]
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    acc0, acc1 = np.load(outs[0]), np.load(outs[1])
    assert np.array_equal(acc0, acc1)  # both ranks hold the reduced frame
    full, feat = _single(spp, features=True)
    assert np.allclose(acc0, full, rtol=2e-6, atol=1e-6)
    assert full[..., 3].min() > 0
    assert np.allclose(np.load(outs[0] + ".feat.npy"), feat, rtol=2e-6, atol=1e-6)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_single_process_device_group():
    """hjk_create(device_ids = {0, 1}): one context, both GPUs; pass p runs on device p mod 2."""
    sys.path.insert(0, ROOT)
    import hijiki_b200 as hj
    import _libs
    ctx = hj.Context([0, 1])
    assert ctx.get_info("n_devices") == 2
    compiled = hj.Scene.from_obj(_libs.CBOX_OBJ, put_cbox_spheres=True).compile()
    ctx.scene_upload(compiled)
    blocks = hj.ImageBlockGenerator(W, H, BS, SPP).blocks()
    ctx.frame_begin(W, H)
    st = ctx.render(blocks, hj.make_params(max_bounces=BOUNCES))
    acc = ctx.readback(normalise=False)
    assert np.array_equal(acc, ctx.readback(normalise=False))
    full, _ = _single(SPP)
    assert st.n_paths == W * H * SPP
    assert np.allclose(acc, full, rtol=2e-6, atol=1e-6)
    img = ctx.readback(normalise=True)
    assert np.isfinite(img).all()
    ctx.close()
