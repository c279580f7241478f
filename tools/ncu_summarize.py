#!/usr/bin/env python
"""Reduce `ncu --set full` reports to the handful of numbers the docs and bench.py quote.

    python tools/ncu_summarize.py NAME=report.ncu-rep [NAME=report.ncu-rep ...] --out profiles/r02/ncu_summary.json --source-hash $(cat gpurun_out/<tag>/source_hash.txt)

For every kernel launch in a report: duration, DRAM bytes read / written (their sum is the `traffic`
of bench.py's roofline object), DRAM throughput as a fraction of ncu's own peak, threads per executed
warp instruction, launch shape.  Also writes the raw csv next to the summary (NAME_raw.csv)."""
import argparse
import csv
import io
import json
import os
import subprocess

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def rows_of(report):
    text = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units = rows[0], rows[1]
    return header, units, rows[2:], text


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+", help="NAME=path.ncu-rep")
    ap.add_argument("--out", required=True)
    ap.add_argument("--source-hash", default=None,
                    help="digest of the native sources the profiled library was built from (jn_source_hash(), written "
                         "by tools/gpu_ci.sh next to the reports); bench.py quotes a capture only for matching code")
    ap.add_argument("--note", default=None, help="the profiled command, recorded next to the hash")
    args = ap.parse_args()
    summary = {}
    if os.path.exists(args.out):
        summary = json.load(open(args.out))
    for spec in args.reports:
        name, path = spec.split("=", 1)
        header, units, rows, text = rows_of(path)
        with open(os.path.join(os.path.dirname(args.out), f"ncu_full_{name}_raw.csv"), "w") as f:
            f.write(text)
        col = {h: i for i, h in enumerate(header)}

        def val(row, key):
            i = col.get(key)
            if i is None or row[i] == "":
                return None
            return float(row[i].replace(",", "")) * UNIT.get(units[i], 1.0)

        kernels = {}
        for row in rows:
            kname = row[col["Kernel Name"]]
            short = kname.split("(")[0].replace("void ", "").replace("jnk::", "").replace("(int)", "").replace("(bool)", "")
            rd, wr = val(row, "dram__bytes_read.sum"), val(row, "dram__bytes_write.sum")
            kernels.setdefault(short, []).append({
                "duration_s": val(row, "gpu__time_duration.sum"),
                "dram_read_bytes": rd, "dram_write_bytes": wr,
                "dram_traffic_bytes": None if rd is None or wr is None else rd + wr,
                "dram_pct_of_ncu_peak": val(row, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "threads_per_inst": val(row, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                "sm_cycles_active_min": val(row, "sm__cycles_active.min"),
                "sm_cycles_active_max": val(row, "sm__cycles_active.max"),
                "registers": val(row, "launch__registers_per_thread"),
                "grid": row[col["Grid Size"]] if "Grid Size" in col else None,
                "block": row[col["Block Size"]] if "Block Size" in col else None,
            })
        summary[name] = kernels
        if args.source_hash:
            summary.setdefault("_meta", {})[name] = {"source_hash": args.source_hash.strip(), "command": args.note}
    with open(args.out, "w") as f:
        json.dump(summary, f, indent=1)
    for name, kernels in summary.items():
        if name == "_meta":
            continue
        for k, recs in kernels.items():
            for r in recs:
                t = r["dram_traffic_bytes"]
                print(f"{name:16s} {k:44s} {1e6 * (r['duration_s'] or 0):9.1f} us  traffic {0 if t is None else t / 1e9:7.3f} GB  "
                      f"dram {r['dram_pct_of_ncu_peak']}%  thr/inst {r['threads_per_inst']}")


if __name__ == "__main__":
    main()
