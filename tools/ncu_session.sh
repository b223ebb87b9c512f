#!/bin/bash
# One ncu session on a B200 box (run through gpurun): launch list of the bench command, full-set captures of part of
# one forward pass (plain launches) and of the selection / mask kernels.  Outputs under gpurun_out/ (reports larger
# than a few MB are exported to CSV and deleted: gpurun_out/ is capped at 64 MiB).
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r1b_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
python tools/profile_once.py > $O/ncu_once_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'conv|stem' -s 70 -c 12 -o $O/r1b_conv_a -f python tools/profile_once.py > $O/ncu_once_conv_a.log 2>&1
ncu -i $O/r1b_conv_a.ncu-rep --page raw --csv > $O/r1b_conv_a_raw.csv 2>/dev/null
rm -f $O/r1b_conv_a.ncu-rep
ncu --set full --clock-control none -k regex:'conv|stem' -s 115 -c 25 -o $O/r1b_conv_b -f python tools/profile_once.py > $O/ncu_once_conv_b.log 2>&1
ncu -i $O/r1b_conv_b.ncu-rep --page raw --csv > $O/r1b_conv_b_raw.csv 2>/dev/null
rm -f $O/r1b_conv_b.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:'mask_decode|nms_kernel|decode_filter8' -s 3 -c 3 -o $O/r1b_tail_full -f python tools/profile_once.py > $O/ncu_once_tail.log 2>&1
ncu -i $O/r1b_tail_full.ncu-rep --page raw --csv > $O/r1b_tail_full_raw.csv 2>/dev/null
ls -la $O/*.ncu-rep
du -sh $O
