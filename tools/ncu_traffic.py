#!/usr/bin/env python
"""Per-kernel time and DRAM traffic of one step from `ncu --set full` captures of both arms ->
profiles/r02_traffic.json (bench.py reads `roofline.traffic` for the dominant kernel from it) and a readable table.

    python tools/ncu_traffic.py ours=gpurun_out/prof/r02_all_ours.ncu-rep reference=gpurun_out/prof/r02_all_ref.ncu-rep
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = [("project_kernel", "project"), ("radix_hist", "sort_hist"), ("radix_onesweep", "sort_pass"), ("DeviceScan", "depth_scan"),
         ("emit_instances", "emit_instances"), ("ranges_and_records", "ranges_records"), ("blend_forward", "blend_forward"),
         ("blend_backward", "blend_backward"), ("geometry_backward", "geometry_backward"), ("finalize_count", "finalize_count"),
         ("preprocessCUDA", "preprocessCUDA"), ("renderCUDA", "renderCUDA"), ("duplicateWithKeys", "duplicateWithKeys"),
         ("identifyTileRanges", "identifyTileRanges"), ("computeCov2DCUDA", "computeCov2DCUDA"), ("DeviceRadixSort", "cub_radix_sort"),
         ("RadixSort", "cub_radix_sort")]


def short(name):
    for pat, lab in STAGE:
        if pat in name:
            return lab
    return name.split("(")[0][-40:]


def parse(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, k):
        v, u = float(r[ix[k]].replace(",", "") or 0), units[ix[k]]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0,
                 "nsecond": 1e-6, "ns": 1e-6, "second": 1e3}.get(u, 1.0)
        return v * scale
    out = []
    for r in rows[2:]:
        out.append(dict(kernel=short(r[ix["Kernel Name"]]), full=r[ix["Kernel Name"]][:100], ms=val(r, "gpu__time_duration.sum"),
                        dram_read=val(r, "dram__bytes_read.sum"), dram_write=val(r, "dram__bytes_write.sum"),
                        dram_pct=0.0,
                        issue_pct=float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]] or 0)))
    for k in out:       # DRAM utilisation against the measured copy bandwidth of MEASURED_PEAKS.json (6572.5 GB/s)
        k["dram_pct"] = 100.0 * (k["dram_read"] + k["dram_write"]) / max(k["ms"] * 1e-3, 1e-12) / 6572.5e9
    return out


def main():
    res, table = {}, []
    outdir = os.path.join(ROOT, "profiles")
    args = []
    for a in sys.argv[1:]:
        if a.startswith("--out="):
            outdir = a[6:]
        else:
            args.append(a)
    for arg in args:
        arm, rep = arg.split("=", 1)
        ks = parse(rep)
        agg = {}
        for k in ks:
            a = agg.setdefault(k["kernel"], dict(launches=0, ms=0.0, dram_bytes=0.0))
            a["launches"] += 1; a["ms"] += k["ms"]; a["dram_bytes"] += k["dram_read"] + k["dram_write"]
        res[arm] = agg
        table.append("## %s (%s)\n" % (arm, rep))
        table.append("%-28s %3s %10s %12s %12s %8s %8s" % ("kernel", "n", "ms", "dram rd MB", "dram wr MB", "dram %", "issue %"))
        for k in ks:
            table.append("%-28s %3d %10.4f %12.2f %12.2f %8.1f %8.1f" % (k["kernel"], 1, k["ms"], k["dram_read"] / 1e6,
                                                                     k["dram_write"] / 1e6, k["dram_pct"], k["issue_pct"]))
        table.append("total: %.4f ms, %.1f MB DRAM traffic (cold-cache, serialised replays: shares, not bench values)\n"
                     % (sum(k["ms"] for k in ks), sum(k["dram_read"] + k["dram_write"] for k in ks) / 1e6))
    json.dump(res, open(os.path.join(outdir, "r02_traffic.json"), "w"), indent=1, sort_keys=True)
    open(os.path.join(outdir, "r02_all_kernels_ncu_full.txt"), "w").write("\n".join(table))
    print("\n".join(table))


if __name__ == "__main__":
    main()
