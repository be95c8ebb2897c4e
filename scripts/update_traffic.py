"""profiles/traffic_k_traverse.json from ncu metric logs of one frame per workload:
    python scripts/update_traffic.py c2=gpurun_out/r2last/frame_c2.csv c4=gpurun_out/r2last/trav_c4.csv
For each workload: mean over the frame's k_traverse launches of dram__bytes_read.sum + dram__bytes_write.sum, the number of
launches, the capture's name and the fingerprint of the kernel source it was taken from (bench.kernel_source_sha: bench.py
quotes the figure only while the source in the tree is the captured one)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from summarize_ncu_table import load  # noqa: E402

out = {"what": "dram__bytes_read.sum + dram__bytes_write.sum per k_traverse launch (mean over the launches of one frame), ncu --clock-control none",
       "workloads": {}}
for arg in sys.argv[1:]:
    key, path = arg.split("=")
    ls = [d for d in load(path) if d["k"].startswith("k_traverse")]
    per = [d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in ls]
    out["workloads"][key] = {"dram_bytes_per_launch": sum(per) / len(per), "launches": len(per), "per_launch": per,
                             "capture": os.path.basename(path) + " (profiles/r2_final_ncu_*.md)", "kernel_src_sha": bench.kernel_source_sha()}
json.dump(out, open(os.path.join(bench.ROOT, "profiles", "traffic_k_traverse.json"), "w"), indent=1)
print(json.dumps({k: (v["dram_bytes_per_launch"], v["launches"]) for k, v in out["workloads"].items()}))
