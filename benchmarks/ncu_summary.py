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


def summarise_all(args, header, units, rows):
    import os
    peak = args.peak_gbs
    if peak is None:
        try:
            peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            peak = 6650.0
    bytes_map = json.load(open(args.bytes_json))["kernels"] if args.bytes_json else {}
    seen = {}
    out = {"command": args.command, "workload": args.workload, "hbm_peak_gbs": peak,
           "note": "ncu replays every launch cold-cache and serialised: gpu__time_duration is an upper bound of the in-pipeline "
                   "time; frac = algorithmic bytes / duration / measured HBM peak; traffic_ratio = dram bytes / algorithmic bytes",
           "launches": []}

    def val(row, m):
        if m not in header:
            return None, None
        i = header.index(m)
        try:
            return float(row[i].replace(",", "")), units[i]
        except ValueError:
            return None, units[i]

    def strip_args(full):  # "ns::k<(bool)0>(T1, T2)" -> "ns::k<(bool)0>"
        full = full.strip()
        if not full.endswith(")"):
            return full
        depth = 0
        for i in range(len(full) - 1, -1, -1):
            depth += full[i] == ")"
            depth -= full[i] == "("
            if depth == 0:
                return full[:i]
        return full

    for row in rows:
        name = strip_args(row[header.index("Kernel Name")]).replace("void ", "").strip()
        base = name.split("<")[0].split("::")[-1]
        seen[base] = seen.get(base, 0) + 1
        e = {"kernel": name, "launch": seen[base]}
        dur, du = val(row, "gpu__time_duration.sum")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(du, 1e-6)
        e["duration_ms"] = dur * scale if dur is not None else None
        r, ru = val(row, "dram__bytes_read.sum")
        w, wu = val(row, "dram__bytes_write.sum")
        if r is not None and w is not None:
            e["dram_bytes"] = r * UNIT_SCALE.get(ru, 1.0) + w * UNIT_SCALE.get(wu, 1.0)
        for m, k in (("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
                     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                     ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps_per_cycle"),
                     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
                     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
                     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct_of_peak"),
                     ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("smsp__inst_executed.sum", "warp_instructions"),
                     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
                     ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts")):
            v, _ = val(row, m)
            if v is not None:
                e[k] = v
        b = bytes_map.get(f"{base}#{seen[base]}") or bytes_map.get(base + ("<pairs>" if "1>" in name or "true>" in name else "")) or bytes_map.get(base)
        if b and e.get("duration_ms"):
            e["workload"] = b["workload"]
            e["algorithmic_bytes"] = b["algorithmic_bytes"]
            e["achieved_GBps"] = b["algorithmic_bytes"] / (e["duration_ms"] * 1e-3) / 1e9
            e["frac"] = e["achieved_GBps"] / peak
            if e.get("dram_bytes"):
                e["traffic_ratio"] = e["dram_bytes"] / b["algorithmic_bytes"]
        out["launches"].append(e)
    json.dump(out, open(args.out, "w"), indent=1)
    for e in out["launches"]:
        print(f"{e['kernel'][:60]:60s} #{e['launch']:<2d} {e.get('duration_ms') or 0:9.3f} ms  frac {e.get('frac') or 0:5.2f}  traffic x{e.get('traffic_ratio') or 0:4.2f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default="")
    ap.add_argument("--out", required=True)
    ap.add_argument("--command", default="")
    ap.add_argument("--workload", default="")
    ap.add_argument("--algorithmic-bytes", type=float, default=None)
    ap.add_argument("--all", action="store_true", help="summarise EVERY launch in the report (one entry per launch, in launch order)")
    ap.add_argument("--bytes-json", default="", help="benchmarks/prof_all.py's kernel -> algorithmic bytes map (--all)")
    ap.add_argument("--peak-gbs", type=float, default=None, help="measured HBM peak for the frac column (default MEASURED_PEAKS.json)")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    if args.all:
        return summarise_all(args, header, units, [r for r in rows[2:] if len(r) == len(header)])
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
