"""Summarise an .ncu-rep (run here, no GPU): per launch the headline metrics, and for one launch the top stall sites.
usage: python tools/ncu_summary.py rep.ncu-rep [launch_index_for_source_view]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "smsp__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]
for n, r in enumerate(data):
    print(f"--- launch {n}")
    for k in keys:
        if k in idx:
            print(f"  {k:75s} {r[idx[k]][:90]} {rows[1][idx[k]]}")
if len(sys.argv) > 2:
    n = int(sys.argv[2])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(n), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = None
    for i, r in enumerate(rows):
        if r and r[0] == "Address":
            h = i
            break
    hdr = rows[h]
    idx = {x: i for i, x in enumerate(hdr)}
    data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0] != "Address"]
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    stalls = [x for x in hdr if x.startswith("stall_") and "Not Issued" not in x]
    agg = {x: sum(int(r[idx[x]]) for r in data) for x in stalls}
    print(f"total samples {tot}; stall mix:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:28]:
        s = {x: int(r[idx[x]]) for x in stalls if int(r[idx[x]]) > 0}
        print(f"{int(r[idx['# Samples']]):6d} {r[idx['Source']].strip()[:64]:64s} {dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])}")
