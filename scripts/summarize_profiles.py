"""Turn ncu outputs brought back in gpurun_out/ into the compact, tracked summaries under profiles/.

    python scripts/summarize_profiles.py launches gpurun_out/<launches>.csv profiles/<name>.md
    python scripts/summarize_profiles.py full     gpurun_out/<report>.ncu-rep profiles/<name>.md [profiles/traffic.json]
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEY_METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occ limit regs (CTAs)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("dv::", "")[:90]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="ignore")) if len(r) > 14 and r[0] != "ID"]
    agg = collections.OrderedDict()
    for r in rows:
        if r[12] != "gpu__time_duration.sum":
            continue
        t = float(r[14].replace(",", ""))
        unit = r[13]
        t_us = t / 1e3 if unit in ("nsecond", "ns") else (t if unit in ("usecond", "us") else t * 1e3)
        k = short(r[4])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t_us
    ours = {k: v for k, v in agg.items() if "_kernel" in k and not k.startswith("at::")}
    tot_ours = sum(v[1] for v in ours.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list: `{src}`\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — cold-cache, serialised launches: compare SHARES, "
                "not absolute times.\n\n| kernel | launches | total us | avg us | share of our kernels |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            share = f"{100 * t / tot_ours:.1f}%" if k in ours and tot_ours else "—"
            f.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {share} |\n")
    print("wrote", dst)


def full(src, dst, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary: `{src}`\n\n(`--clock-control none --import-source on`; one row block per profiled launch)\n\n")
        for r in rows[2:]:
            name = short(r[idx["Kernel Name"]])
            f.write(f"## `{name}`\n\n| metric | value |\n|---|---|\n")
            vals = {}
            for m, label in KEY_METRICS:
                if m in idx:
                    vals[m] = r[idx[m]]
                    f.write(f"| {label} (`{m}`) | {r[idx[m]]} {units[idx[m]]} |\n")
            f.write("\n")

            def to_bytes(m):
                v = float(vals[m].replace(",", ""))
                u = units[idx[m]]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
            try:
                key = name.split("<")[0]
                traffic.setdefault(key, []).append(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
            except Exception:
                pass
    if traffic_json:
        json.dump({k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_profiled": len(v)} for k, v in traffic.items()},
                  open(traffic_json, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
