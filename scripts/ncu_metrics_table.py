"""Pivot an `ncu --csv --metrics ...` log into one row per launch (short metric names)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = next(i for i, r in enumerate(rows) if "Metric Name" in r)
h = rows[hi]
iid, ik, im, iv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
launches = {}
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    d = launches.setdefault(int(r[iid]), {"k": r[ik].split("(")[0][-22:]})
    d[r[im]] = float(r[iv].replace(",", ""))
short = [("gpu__time_duration.sum", "us", 1e-3), ("smsp__inst_executed.sum", "Minst", 1e-6), ("lanes", "lanes", 1), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
         ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1), ("l1tex__t_sector_hit_rate.pct", "L1%", 1), ("lts__t_sector_hit_rate.pct", "L2%", 1),
         ("dram", "dramMB", 1e-6), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%", 1), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%", 1),
         ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%", 1)]
stalls = ["long_scoreboard", "short_scoreboard", "wait", "not_selected", "math_pipe_throttle", "branch_resolving", "mio_throttle", "no_instruction", "dispatch_stall", "lg_throttle", "barrier"]
print(" id kernel                 " + " ".join(f"{n:>8s}" for _, n, _ in short) + "  stalls/issue: " + " ".join(s[:6] for s in stalls))
tot = {}
for i in sorted(launches):
    d = launches[i]
    d["lanes"] = d.get("smsp__thread_inst_executed.sum", 0) / max(1.0, d.get("smsp__inst_executed.sum", 1))
    d["dram"] = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    print(f"{i:3d} {d['k']:22s} " + " ".join(f"{d.get(m, 0) * sc:8.1f}" for m, _, sc in short) + "   " +
          " ".join(f"{d.get('smsp__average_warps_issue_stalled_' + s + '_per_issue_active.ratio', 0):6.2f}" for s in stalls))
    for m in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum"):
        tot[m] = tot.get(m, 0) + d.get(m, 0)
print(f"sum: {tot.get('gpu__time_duration.sum', 0) * 1e-3:.1f} us, {tot.get('smsp__inst_executed.sum', 0) * 1e-6:.1f} M warp instructions, "
      f"{tot.get('smsp__thread_inst_executed.sum', 0) / max(1.0, tot.get('smsp__inst_executed.sum', 1)):.2f} lanes")
