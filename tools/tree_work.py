"""Node steps and primitive tests per ray of the host-built wide BVH, on the rays of a real render (all bounces),
counted on the CPU by the host-compiled test harness (tests/native, ht_work_stats): the two quantities the trace
kernel's time is made of.  No GPU needed: a tree change can be judged here before a GPU run confirms it.

    python tools/tree_work.py [cbox|cbox_spheres|terrain:N|spheres:N] [width height spp max_bounces]
Environment knobs of the builder (HJK_BVH_PRIM_COST, ...) apply."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import _libs


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "cbox"
    w, h, spp, mb = (int(a) for a in sys.argv[2:6]) if len(sys.argv) > 5 else (256, 144, 2, 8)
    L = _libs.hosttest()
    L.ht_work_stats.restype = C.c_int
    L.ht_work_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(_libs._abi.HjkParams), C.c_void_p]
    if kind.startswith("terrain"):
        scene = _libs.HostScene.terrain(L, int(kind.split(":")[1]))
    elif kind.startswith("spheres"):
        scene = _libs.HostScene.spheres(L, int(kind.split(":")[1]))
    else:
        scene = _libs.HostScene.from_obj(L, put_spheres=kind == "cbox_spheres")
    err = C.create_string_buffer(256)
    hh = L.ht_create(C.byref(scene.view), 1e-5, err, 256)
    assert hh, err.value
    n_nodes, n_prims, depth, sah, pad, ms = C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_float(), C.c_float(), C.c_int()
    L.ht_bvh_stats(hh, C.byref(n_nodes), C.byref(n_prims), C.byref(depth), C.byref(sah), C.byref(pad), C.byref(ms))
    blocks = _libs.generate_blocks(L, w, h, spp)
    prm = _libs.hjk_params(max_bounces=mb)
    out = np.zeros(6, dtype=np.uint64)
    rc = L.ht_work_stats(hh, _libs.ptr(blocks), blocks.size, C.byref(prm), _libs.ptr(out))
    assert rc == 0
    ce, cn, cp, ae, an, ap = (float(v) for v in out)
    print(f"{kind}: {n_nodes.value} nodes, {n_prims.value} prims, depth {depth.value}, SAH {sah.value:.3f}")
    print(f"  closest-hit rays {int(ce)}: {cn / ce:.3f} node steps, {cp / ce:.3f} primitive tests per ray")
    if ae:
        print(f"  any-hit rays     {int(ae)}: {an / ae:.3f} node steps, {ap / ae:.3f} primitive tests per ray")
    tot = ce + ae
    print(f"  all rays: {(cn + an) / tot:.3f} node steps, {(cp + ap) / tot:.3f} primitive tests per ray")


if __name__ == "__main__":
    main()
