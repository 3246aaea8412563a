"""Fold the raw page of an ncu --set full capture into profiles/ncu_trace_metrics.json — the profiler-only figures
bench.py reports beside its live measurements (DRAM bytes per traced ray, issue-active cycles and lanes per
instruction of the trace kernel) — and a selected-columns CSV for profiles/.

    ncu -i capture.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_trace_metrics.py raw.csv --workload cbox --rays 216266801 --source profiles/r02c_... \
        [--selected profiles/r02c_ncu_full_wave_selected.csv]

`--rays` = rays traced by the captured launches (extension + shadow of the captured wave, from the bench line the
captured command printed).  The summary records the digest of the device sources, so bench.py can tell when it
has gone stale."""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

COLS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu.sum"]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
              "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("--workload", required=True)
    ap.add_argument("--rays", type=float, required=True)
    ap.add_argument("--source", required=True, help="name of the committed capture summary these come from")
    ap.add_argument("--selected", help="also write the selected columns of every launch here")
    ap.add_argument("--kernel", default="k_trace_coop")
    args = ap.parse_args()
    with open(args.raw_csv) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.reader(lines))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = {name: i for i, name in enumerate(head)}
    cols = [c for c in COLS if c in idx]

    def val(row, name):
        raw = row[idx[name]].replace(",", "")
        try:
            v = float(raw)
        except ValueError:
            return None
        return v * UNIT_SCALE.get(units[idx[name]], 1.0)

    if args.selected:
        with open(args.selected, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["ID", "Kernel Name"] + cols)
            w.writerow(["", ""] + [units[idx[c]] for c in cols])
            for r in body:
                w.writerow([r[idx["ID"]], r[idx["Kernel Name"]]] + [r[idx[c]] for c in cols])
    sel = [r for r in body if args.kernel in r[idx["Kernel Name"]]]
    if not sel:
        raise SystemExit(f"no launch of {args.kernel} in {args.raw_csv}")
    t = [val(r, "gpu__time_duration.sum") for r in sel]
    inst = [val(r, "smsp__inst_executed.sum") for r in sel]
    dram = sum(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in sel)
    issue = sum(val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") * ti for r, ti in zip(sel, t)) / sum(t)
    lanes = sum(val(r, "smsp__thread_inst_executed_per_inst_executed.ratio") * ii for r, ii in zip(sel, inst)) / sum(inst)
    warps = sum(val(r, "sm__warps_active.avg.pct_of_peak_sustained_active") * ti for r, ti in zip(sel, t)) / sum(t)
    from bench import NCU_METRICS, kernel_source_sha
    try:
        with open(NCU_METRICS) as f:
            out = json.load(f)
    except Exception:
        out = {"workloads": {}}
    sha = kernel_source_sha()
    if out.get("kernel_source_sha") != sha:
        out = {"workloads": {}}  # figures of other sources do not mix
    out["kernel_source_sha"] = sha
    out["source"] = args.source
    out["workloads"][args.workload] = {
        "kernel": args.kernel, "launches": len(sel), "rays": args.rays, "dram_bytes": dram,
        "dram_bytes_per_ray": dram / args.rays, "issue_active_pct": issue, "lanes_per_instruction": lanes,
        "warps_active_pct": warps, "time_ms_under_ncu": sum(t) * 1e3, "capture": args.source}
    with open(NCU_METRICS, "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print(json.dumps(out["workloads"][args.workload], indent=1))


if __name__ == "__main__":
    main()
