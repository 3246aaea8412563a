"""N > 1 host logic on CPU (gloo, world_size 2): the sample-pass split of SURVEY.md §8e.

Each rank takes the passes p = rank (mod world) of one block list, integrates + reconstructs them
(with the host-compiled kernel logic of tests/native — test infrastructure, the product's
multi-GPU path is the same split feeding hjk_render on each GPU), and one all-reduce(sum) of the
accumulators must reproduce the single-rank frame up to fp32 summation order.  bench.py's
step_slice is checked the same way."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import _libs
    import hijiki_b200 as hj
    from bench import step_slice

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    H = _libs.hosttest()
    scene = _libs.HostScene.from_obj(H, _libs.CBOX_OBJ, False, False)
    err = C.create_string_buffer(256)
    h = H.ht_create(C.byref(scene.view), 1e-5, err, 256)
    w, hgt, bs, spp = 96, 64, 64, 4
    blocks = _libs.generate_blocks(H, w, hgt, spp, block_size=bs)
    bpp = 2
    mine = hj.split_passes(blocks, bpp, rank, world)
    assert mine.size == blocks.size // world
    # bench.py's per-step slice: step 0 with spp_per_step = spp / world covers the same passes
    sl = step_slice(blocks, bpp, 0, rank, world, spp // world)
    assert sorted(sl["id"].tolist()) == sorted(mine["id"].tolist())
    acc = np.zeros((hgt, w, 4), np.float32)
    hp = _libs.hjk_params(max_bounces=6)
    assert H.ht_render(h, _libs.ptr(mine), mine.size, C.byref(hp), _libs.ptr(acc), None, None) == 0
    t = torch.from_numpy(acc)
    dist.all_reduce(t)  # the one collective of the path
    if rank == 0:
        full = np.zeros((hgt, w, 4), np.float32)
        assert H.ht_render(h, _libs.ptr(blocks), blocks.size, C.byref(hp), _libs.ptr(full), None, None) == 0
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
        np.save(os.path.join(out_dir, "full.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


def test_sample_pass_split_allreduce_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    full = np.load(tmp_path / "full.npy")
    assert np.isfinite(full).all() and full[..., 3].min() > 0
    assert np.allclose(reduced, full, rtol=2e-6, atol=1e-6)


def test_reference_arm_exits_quietly_on_nonzero_rank():
    """bench.py --impl reference: under torchrun only rank 0 works and prints."""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--gpus", "2"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_prints_one_contract_line():
    """bench.py --impl reference on rank 0: exactly ONE line on stdout, the JSON of the bench contract with the
    reference-arm keys (the CPU oracle timed on the host cores on a bounded sample)."""
    import json
    import subprocess
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--cpu-seconds", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s (all bounces)" and d["unit"] == "Mrays/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_never_maps_the_product_library():
    import subprocess
    """bench.py --impl reference builds its scene, BVH2 and block list with the tests-side host library and traces
    with the oracle: libhijiki_b200.so must not be in the process at all."""
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-seconds', '1']\n"
            "runpy.run_path('bench.py', run_name='__main__')\n"
            "maps = open('/proc/self/maps').read()\n"
            "sys.stderr.write('MAPPED:' + ','.join(sorted({l.split()[-1] for l in maps.splitlines() if '.so' in l and "
            "('hijiki' in l or 'oracle' in l or 'hosttest' in l)})))\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    mapped = r.stderr.split("MAPPED:")[-1]
    assert "liboracle.so" in mapped and "libhjk_hosttest.so" in mapped
    assert "libhijiki_b200" not in mapped, mapped
