"""ncu `--csv --metrics ...` log -> a markdown table for profiles/.
    python scripts/summarize_ncu_table.py launches <csv> <out.md> "<title>"     one row per launch
    python scripts/summarize_ncu_table.py kernels  <csv> <out.md> "<title>"     one row per kernel (aggregated)
DRAM GB/s = (dram__bytes_read + dram__bytes_write) / gpu__time_duration; HBM peak from MEASURED_PEAKS.json (6583.5 GB/s)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6583.5


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Metric Name" in r)
    h = rows[hi]
    iid, ik, im, iv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    launches = {}
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        name = r[ik].split("(")[0].replace("void ", "").replace("crt::", "")
        d = launches.setdefault(int(r[iid]), {"k": name})
        try:
            d[r[im]] = float(r[iv].replace(",", ""))
        except ValueError:
            pass
    return [launches[i] for i in sorted(launches)]


def derived(d):
    t_ns = d.get("gpu__time_duration.sum", 0.0)
    dram = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    inst = max(1.0, d.get("smsp__inst_executed.sum", 1.0))
    return {"us": t_ns * 1e-3, "lanes": d.get("smsp__thread_inst_executed.sum", 0.0) / inst, "minst": inst * 1e-6,
            "issue": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0),
            "occ": d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0), "regs": d.get("launch__registers_per_thread", 0.0),
            "rd": d.get("dram__bytes_read.sum", 0.0) * 1e-6, "wr": d.get("dram__bytes_write.sum", 0.0) * 1e-6,
            "gbs": dram / max(1.0, t_ns), "l1": d.get("l1tex__t_sector_hit_rate.pct", 0.0), "l2": d.get("lts__t_sector_hit_rate.pct", 0.0),
            "stall": {k.split("stalled_")[1].split("_per_")[0]: v for k, v in d.items() if "issue_stalled" in k}}


def top_stalls(st, n=3):
    return ", ".join(f"{k.replace('_', ' ')} {v:.1f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:n])


def main():
    mode, src, dst, title = sys.argv[1:5]
    ls = load(src)
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nSource: `{src}` (ncu `--metrics ... --clock-control none`, a few replays per launch; times under ncu are serialised and "
                f"cold-cache: compare shares, not absolutes). Warp exec eff = smsp__thread_inst_executed / smsp__inst_executed (active lanes per "
                f"instruction, of 32). DRAM GB/s = (dram__bytes_read + dram__bytes_write) / gpu__time_duration; HBM peak measured on this pool: "
                f"{PEAK:.1f} GB/s. Stalls: warps stalled per issued instruction, three largest reasons.\n\n")
        if mode == "launches":
            f.write("| # | kernel | time us | warp exec eff | issue active % | warps active % | regs | DRAM read MB | DRAM write MB | DRAM GB/s | % of HBM peak | L1 hit % | L2 hit % | warp instr M | top stalls |\n")
            f.write("|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
            for i, d in enumerate(ls):
                v = derived(d)
                f.write(f"| {i} | `{d['k']}` | {v['us']:.1f} | {v['lanes']:.1f} | {v['issue']:.1f} | {v['occ']:.1f} | {v['regs']:.0f} | {v['rd']:.1f} | {v['wr']:.1f} | "
                        f"{v['gbs']:.0f} | {100 * v['gbs'] / PEAK:.1f} | {v['l1']:.1f} | {v['l2']:.1f} | {v['minst']:.1f} | {top_stalls(v['stall'])} |\n")
            tot = sum(derived(d)["us"] for d in ls)
            by = {}
            for d in ls:
                by[d["k"]] = by.get(d["k"], 0.0) + derived(d)["us"]
            f.write("\nShares of the frame under ncu: " + ", ".join(f"`{k}` {100 * v / tot:.1f} %" for k, v in sorted(by.items(), key=lambda x: -x[1])) + ".\n")
        else:
            agg = {}
            for d in ls:
                v = derived(d)
                a = agg.setdefault(d["k"], {"n": 0, "us": 0.0, "thr": 0.0, "inst": 0.0, "issue_t": 0.0, "occ_t": 0.0, "bytes": 0.0, "regs": v["regs"], "stall": {}})
                a["n"] += 1
                a["us"] += v["us"]
                a["thr"] += d.get("smsp__thread_inst_executed.sum", 0.0)
                a["inst"] += d.get("smsp__inst_executed.sum", 0.0)
                a["issue_t"] += v["issue"] * v["us"]
                a["occ_t"] += v["occ"] * v["us"]
                a["bytes"] += (v["rd"] + v["wr"]) * 1e6
                for k, s in v["stall"].items():
                    a["stall"][k] = a["stall"].get(k, 0.0) + s * v["us"]
            tot = sum(a["us"] for a in agg.values())
            f.write("| kernel | launches | total us | share | mean us | warp exec eff | issue active % (time-weighted) | warps active % | regs | DRAM GB/s | % of HBM peak | top stalls |\n")
            f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
            for k, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
                gbs = a["bytes"] / max(1.0, a["us"] * 1e3)
                st = {s: v / max(1e-9, a["us"]) for s, v in a["stall"].items()}
                f.write(f"| `{k}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f} % | {a['us'] / a['n']:.1f} | {a['thr'] / max(1.0, a['inst']):.1f} | "
                        f"{a['issue_t'] / max(1e-9, a['us']):.1f} | {a['occ_t'] / max(1e-9, a['us']):.1f} | {a['regs']:.0f} | {gbs:.0f} | {100 * gbs / PEAK:.1f} | {top_stalls(st)} |\n")
    print(open(dst).read()[:2500])


if __name__ == "__main__":
    main()
