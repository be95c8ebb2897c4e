#!/usr/bin/env python
"""Turn ncu artefacts from gpurun_out/ into small, committed summaries under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r1b.csv profiles/r1_launches_b.md
  python scripts/summarize_ncu.py report   gpurun_out/prof_r1b.ncu-rep   profiles/r1_ncu_full_b.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list `{src}` (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / v[0] / 1e3:.1f} | {v[1] / tot:.3f} |\n")
    print(open(dst).read())


def report(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full capture `{src}`\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
            f.write(f"## launch id {r[hdr.index('ID')]}: `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
