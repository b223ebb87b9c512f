#!/bin/bash
# DRAM bytes and duration of every kernel of ONE forward pass (+ masks), plain launches: the `traffic` figure of
# bench.py's roofline object.  Cheap (three metrics, no --set full).  Output: gpurun_out/r2_dram_per_launch.csv (both passes of
# profile_once.py; tools/ncu_traffic_json.py keeps the last one)
cd "$(dirname "$0")/.."
O=gpurun_out
python tools/profile_once.py > $O/ncu_traffic_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --csv --log-file $O/r2_dram_per_launch.csv python tools/profile_once.py > $O/ncu_traffic.log 2>&1
tail -2 $O/ncu_traffic.log; wc -l $O/r2_dram_per_launch.csv
