#!/usr/bin/env python
"""ncu_summary.py REPORT.ncu-rep OUT.json -- keep the judged numbers of one `ncu --set full` capture.

Reads the raw page of the report (ncu -i ... --page raw --csv) and writes, per kernel launch in
the report, duration, DRAM bytes (read + written = `traffic`), instruction counts, issue
utilisation and the top warp-stall reasons.  The .ncu-rep itself stays in gpurun_out/ (scratch)."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp64.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        k = {"kernel": r[hdr.index("Kernel Name")], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for name in KEEP:
            if name in hdr:
                i = hdr.index(name)
                try:
                    k[name] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
                except ValueError:
                    k[name] = {"value": r[i], "unit": units[i]}
        stalls = []
        for i, name in enumerate(hdr):
            if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), name[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        k["top_stalls_warps_per_issue"] = [{"reason": n, "ratio": v} for v, n in stalls[:6]]

        def scaled(name):
            v = k.get(name)
            if not v or not isinstance(v["value"], float):
                return None
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(v["unit"], 1.0)
            return v["value"] * mult
        rd, wr = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            k["traffic_bytes"] = rd + wr
        kernels.append(k)
    json.dump({"report": rep, "kernels": kernels}, open(out, "w"), indent=1)
    for k in kernels:
        print(k["kernel"][:70], k.get("gpu__time_duration.sum"), "traffic", k.get("traffic_bytes"))


if __name__ == "__main__":
    main()
