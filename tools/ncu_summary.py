#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page raw --csv` into the handful of metrics DESIGN.md cites."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum",
        "lts__t_sector_hit_rate.pct"]


def main(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("per_warp_active.pct")]
    lines = ["# ncu --set full summary of %s" % rep, ""]
    for r in rows[2:]:
        lines.append("## %s  (launch id %s)" % (r[idx["Kernel Name"]][:90], r[idx["ID"]]))
        for k in KEYS:
            if k in idx:
                lines.append("%-72s %s %s" % (k, r[idx[k]], units[idx[k]]))
        top = sorted(((float(r[idx[k]] or 0), k) for k in stalls), reverse=True)[:6]
        for v, k in top:
            lines.append("stall %-66s %.1f %%" % (k.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", ""), v))
        lines.append("")
    text = "\n".join(lines)
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
