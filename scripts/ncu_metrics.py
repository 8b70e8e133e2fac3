"""Print selected raw metrics per profiled launch of an .ncu-rep:  python scripts/ncu_metrics.py rep [kernel-substr] [metric-substr ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
ksub = sys.argv[2] if len(sys.argv) > 2 else ""
msubs = sys.argv[3:] or ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe", "dram__bytes", "smsp__issue_active.avg.pct",
                         "sm__warps_active.avg.pct", "launch__registers", "launch__occupancy_limit", "smsp__average_warps_issue_stalled"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
seen = set()
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if ksub not in name or name in seen:
        continue
    seen.add(name)
    print("==", name[:100])
    for i, h in enumerate(hdr):
        if any(m in h for m in msubs) and r[i] not in ("", "0"):
            print(f"   {h:95s} {r[i]} {units[i]}")
