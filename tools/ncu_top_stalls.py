"""Print the hottest SASS lines (by warp-stall samples) of one kernel launch in an .ncu-rep.
Usage: python tools/ncu_top_stalls.py report.ncu-rep <launch-id> [n]"""
import csv
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(out.splitlines())]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
ix = {h: i for i, h in enumerate(hdr)}
samp = lambda r: int(r[ix["# Samples"]]) if r[ix["# Samples"]].isdigit() else 0
tot = sum(samp(r) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in data if r[ix[h]].isdigit()) for h in stalls}
print("total samples", tot, "| stall mix:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for r in sorted(data, key=lambda r: -samp(r))[:n]:
    st = sorted(((h[6:], int(r[ix[h]])) for h in stalls if r[ix[h]].isdigit() and int(r[ix[h]]) > 0), key=lambda x: -x[1])[:3]
    print(f"{100 * samp(r) / max(tot, 1):5.1f}%  exec {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:64]:64s} {st}")
