"""Per-source-line instruction shares of one kernel from an .ncu-rep captured with --import-source on (-lineinfo build).
    python tools/ncu_lines.py report.ncu-rep [top_n]
Runs `ncu -i ... --page source --csv --print-source cuda,sass` and aggregates "Instructions Executed" / "# Samples"."""
import csv
import io
import subprocess
import sys


def main(argv):
    rep, top = argv[1], int(argv[2]) if len(argv) > 2 else 30
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    for key in ("Duration", "Compute (SM) Throughput", "Issue Slots Busy", "Executed Ipc Elapsed", "L1/TEX Cache Throughput",
                "L2 Cache Throughput", "DRAM Throughput", "Achieved Occupancy", "Registers Per Thread"):
        for line in det.splitlines():
            if line.strip().startswith(key):
                print(" ".join(line.split()))
                break
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    hdr, cur, per, tot = None, None, {}, 0
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr, i_i, i_s = r, r.index("Instructions Executed"), r.index("# Samples")
        elif hdr and len(r) > i_i and r[2] == "-" and r[0].isdigit():
            try:
                n, s = int(r[i_i]), int(r[i_s])
            except ValueError:
                continue
            per[(cur, int(r[0]))] = (n, s, r[1].strip())
            tot += n
    print(f"total warp instructions {tot}")
    for (f, l), (n, s, src) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f[:20]:20s} {l:5d} {100 * n / tot:5.1f} %  samples {s:6d}  {src[:110]}")


if __name__ == "__main__":
    main(sys.argv)
