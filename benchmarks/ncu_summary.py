#!/usr/bin/env python
"""Turns an .ncu-rep capture into the small JSON summary committed under profiles/ (needs only `ncu -i`, no GPU):
    python benchmarks/ncu_summary.py gpurun_out/prof.ncu-rep --kernel convert_tiles --out profiles/x.json \\
        --command "<the ncu command line>" --workload "<what ran>" --algorithmic-bytes 5.5e9
"""
import argparse
import csv
import io
import json
import subprocess

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default="")
    ap.add_argument("--out", required=True)
    ap.add_argument("--command", default="")
    ap.add_argument("--workload", default="")
    ap.add_argument("--algorithmic-bytes", type=float, default=None)
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    launches = [r for r in rows[2:] if len(r) == len(header) and args.kernel in r[header.index("Kernel Name")]]
    if not launches:
        raise SystemExit(f"no launch of a kernel matching '{args.kernel}' in {args.report}")
    row = launches[-1]
    out = {"kernel": row[header.index("Kernel Name")].split("(")[0], "command": args.command, "workload": args.workload, "metrics": {}}
    for m in METRICS:
        if m in header:
            i = header.index(m)
            out["metrics"][m] = {"value": row[i], "unit": units[i]}

    def nbytes(m):
        e = out["metrics"].get(m)
        return float(e["value"].replace(",", "")) * UNIT_SCALE.get(e["unit"], 1.0) if e else None

    r, w = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
    if r is not None and w is not None:
        out["dram_bytes_per_launch"] = r + w
    if args.algorithmic_bytes:
        out["algorithmic_bytes_per_launch"] = args.algorithmic_bytes
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "metrics"}))


if __name__ == "__main__":
    main()
