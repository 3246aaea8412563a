"""Two GPUs, two processes, the LIBRARY's own communicator (hjk_comm_unique_id / hjk_comm_init, NCCL
loaded with dlopen — no torch in the workers): each rank renders the passes p = rank (mod 2) and
hjk_readback all-reduces the accumulator; the result equals the single-GPU frame up to fp32
summation order.  Skipped unless two CUDA devices are visible."""
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


def _worker(rank, world, conn, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import hijiki_b200 as hj
    import _libs
    if rank == 0:
        uid = hj.comm_unique_id()
        conn.send(uid)
    else:
        uid = conn.recv()
    scene = hj.Scene.from_obj(_libs.CBOX_OBJ)
    gen = hj.ImageBlockGenerator(W, H, BS, SPP)
    r = hj.Renderer.new(scene, gen, 128, False, device=rank, max_bounces=BOUNCES, rank=rank, world=world)
    r.ctx.comm_init(uid, rank, world)
    r.render()
    acc = r.ctx.readback(normalise=False)  # all-reduces first
    np.save(out_path, acc)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_library_communicator_allreduce(tmp_path):
    ctx = mp.get_context("spawn")
    a, b = ctx.Pipe()
    outs = [str(tmp_path / f"acc{r}.npy") for r in range(2)]
    procs = [ctx.Process(target=_worker, args=(r, 2, a if r == 0 else b, outs[r])) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    acc0, acc1 = np.load(outs[0]), np.load(outs[1])
    assert np.array_equal(acc0, acc1)  # both ranks hold the reduced frame
    sys.path.insert(0, ROOT)
    import hijiki_b200 as hj
    import _libs
    single = hj.Renderer.new(hj.Scene.from_obj(_libs.CBOX_OBJ), hj.ImageBlockGenerator(W, H, BS, SPP), 128, False,
                             max_bounces=BOUNCES)
    single.render()
    full = single.ctx.readback(normalise=False)
    assert np.allclose(acc0, full, rtol=2e-6, atol=1e-6)
    assert full[..., 3].min() > 0
