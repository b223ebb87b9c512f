"""profiles/ncu_traffic.json from the per-launch DRAM capture of tools/ncu_traffic.sh.
    python tools/ncu_traffic_json.py profiles/r1d_dram_per_launch.csv [workload]"""
import csv, json, os, re, sys

src = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "yolov8s-seg-640-b64"
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
launch = {}
for r in rows:
    d = launch.setdefault(r[ii], {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
# a capture of several forward passes: keep the last one (it starts at the last stem launch)
order = sorted(launch, key=lambda k: int(k))
stems = [i for i, k in enumerate(order) if "stem_tc_kernel" in launch[k]["name"]]
if stems:
    order = order[stems[-1]:]
per, conv_b, conv_us, all_us, n_conv = {}, 0.0, 0.0, 0.0, 0
for d in (launch[k] for k in order):
    name = re.sub(r"^void ", "", d["name"]).split("(")[0]
    k = per.setdefault(name, {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "us_under_ncu": 0.0})
    k["launches"] += 1
    k["dram_read_bytes"] += d.get("dram__bytes_read.sum", 0.0)
    k["dram_write_bytes"] += d.get("dram__bytes_write.sum", 0.0)
    us = d.get("gpu__time_duration.sum", 0.0) / 1e3
    k["us_under_ncu"] += us
    all_us += us
    if any(k in name for k in ("conv_tc2_kernel", "conv3_halo_kernel", "conv3_halo2_kernel", "conv_tc2p_kernel")):
        conv_b += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        conv_us += us
        n_conv += 1
out = {"workload": workload,
       "source": f"{src} (ncu dram__bytes_read/write.sum, one forward pass, plain launches, tools/ncu_traffic.sh)",
       "conv_launches": n_conv, "conv_dram_bytes_per_step": conv_b, "conv_us_under_ncu": conv_us,
       "all_kernels_us_under_ncu": all_us, "conv_share_of_step_under_ncu": conv_us / all_us if all_us else None,
       "per_kernel": per}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(path, "conv launches", n_conv, "conv GB", conv_b / 1e9, "share", out["conv_share_of_step_under_ncu"])
