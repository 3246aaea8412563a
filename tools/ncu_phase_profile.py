"""Fold the source page of an ncu capture of k_trace_coop into phases of the traversal loop.

    ncu -i capture.ncu-rep --page source --csv --print-source cuda,sass --launch-skip N --launch-count 1 > src.csv
    python tools/ncu_phase_profile.py src.csv

Prints, per phase: share of executed warp instructions, share of stall samples, lanes per instruction.
Instructions inlined from hjk_math.cuh inherit the phase of the preceding instruction in address order.
The capture must come from a build of the same sources (-lineinfo): phases are located by the marker
comments of traverse_queue_coop and the function heads of traverse.cuh.
"""
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_of(path, pattern, start=0):
    for i, ln in enumerate(open(path).read().splitlines(), 1):
        if i > start and re.search(pattern, ln):
            return i
    raise SystemExit(f"marker {pattern!r} not found in {path}")


def phase_tables():
    k = os.path.join(ROOT, "hijiki_b200/csrc/device/kernels.cuh")
    t = os.path.join(ROOT, "hijiki_b200/csrc/device/traverse.cuh")
    coop = line_of(k, r"void traverse_queue_coop\(")
    marks = [("refill", line_of(k, r"// ---- refill", coop)), ("node_step", line_of(k, r"// ---- node step", coop)),
             ("prim_decide+per_lane", line_of(k, r"// ---- primitive tests", coop)),
             ("pooled", line_of(k, r"const uint32_t pooled = cnt;", coop)),
             ("advance", line_of(k, r"// ---- advance / finish", coop)),
             ("kernel_entry", line_of(k, r"^// GUARD: the scene contains spheres", coop))]
    io = [("io_load/store", line_of(k, r"^struct WaveIO"), line_of(k, r"^struct BatchIO"))]
    stack = [("stack", line_of(k, r"^struct DevStack"), line_of(k, r"^// Source of rays"))]
    kt = [(n, a, b) for (n, a), (_, b) in zip(marks, marks[1:])] + io + stack
    heads = [("trav_init", line_of(t, r"void trav_init\(")), ("intersect_node", line_of(t, r"uint32_t intersect_node\(")),
             ("intersect_prim", line_of(t, r"^HJK_HD bool intersect_prim\(.*\n?")), ("tie_mode", line_of(t, r"^// EXACT-TIE MODE"))]
    tt = [(n, a, b) for (n, a), (_, b) in zip(heads, heads[1:])]
    return {"kernels.cuh": kt, "traverse.cuh": tt}


def main():
    tables = phase_tables()
    rows = list(csv.reader(open(sys.argv[1])))
    recs, f, hdr, cur = {}, None, None, 0
    for r in rows:
        if r and r[0] == "File Path":
            f = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr is None or len(r) < 10 or (r and r[0] == "Function Name"):
            continue
        elif r[0] != "":
            try:
                cur = int(r[0])
            except ValueError:
                pass
        elif r[2].startswith("0x") and int(r[2], 16) not in recs:
            g = lambda name: float(r[hdr.index(name)] or 0)
            recs[int(r[2], 16)] = (f, cur, g("Instructions Executed"), g("# Samples"), g("Thread Instructions Executed"))
    agg, last = {}, "other"
    for addr in sorted(recs):
        f, line, inst, samp, thr = recs[addr]
        ph = None
        for name, a, b in tables.get(f, []):
            if a <= line < b:
                ph = name
        if ph is None and f in tables:
            ph = "other"
        if ph is None:
            ph = last  # inlined math: phase of the code around it
        last = ph
        a = agg.setdefault(ph, [0.0, 0.0, 0.0])
        a[0] += inst
        a[1] += samp
        a[2] += thr
    ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    for ph, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{ph:22s} instructions {a[0] / ti * 100:5.1f} %   stall samples {a[1] / ts * 100:5.1f} %   "
              f"lanes/instr {a[2] / max(a[0], 1):5.1f}")
    print(f"warp instructions {ti:.0f}")


if __name__ == "__main__":
    main()
