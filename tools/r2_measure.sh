#!/bin/bash
# Round-2 measurement pass on ONE B200 (through gpurun): bench lines and per-op tables of every workload, the ncu launch
# list of the bench command, per-launch DRAM traffic of one forward pass, and full-set captures of the CTA-pair conv
# kernels and the tail kernels (digested into profiles/NCU_SUMMARY.md by tools/ncu_digest.py).
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2_bench_lines.jsonl
python bench.py --dump-ops $O/r2_ops_yolov8s-seg-640-b64.csv > $O/r2_bench_default.log 2> $O/r2_bench_default.err
tail -1 $O/r2_bench_default.log >> $O/r2_bench_lines.jsonl
for w in yolov10n-640-b32 yolov8n-seg-640-b64 yolov8m-seg-1080p-b16 yolov8x-seg-640-b32 yolo11n-seg-640-b64 yolo11s-seg-640-b64 yolo11x-seg-640-b32; do
  python bench.py --workload $w --steps 40 --warmup 3 --no-cpu-baseline --dump-ops $O/r2_ops_$w.csv > $O/r2_bench_$w.log 2>/dev/null
  tail -1 $O/r2_bench_$w.log >> $O/r2_bench_lines.jsonl
done
python - <<PY
import json
for l in open("$O/r2_bench_lines.jsonl"):
    d = json.loads(l)
    print(d["config"]["workload"], round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["pinned_frames"]["value"]),
          "frac", round(d["roofline"]["frac"], 3), "b1", round(d["p50_frame_latency_ms_b1"], 3), round(d["p50_predict_call_ms_b1"], 3), d["clocks"]["sm_mhz"])
PY
# launch list of the bench command
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_ncu_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_ncu_bench.log 2>&1
# per-launch DRAM traffic of one forward pass (plain launches: second pass of profile_once.py)
python tools/profile_once.py > $O/r2_once_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 78 -c 78 \
    --csv --log-file $O/r2_dram_per_launch.csv python tools/profile_once.py > $O/r2_ncu_traffic.log 2>&1
# full-set captures: the CTA-pair kernels (all of one pass), then the tail kernels
ncu --set full --clock-control none --import-source on -k regex:'halo2|tc2p' -s 30 -c 24 -o $O/r2_pair -f python tools/profile_once.py > $O/r2_ncu_pair.log 2>&1
ncu -i $O/r2_pair.ncu-rep --page raw --csv > $O/r2_ncu_full_pair_raw.csv 2>/dev/null
rm -f $O/r2_pair.ncu-rep
ncu --set full --clock-control none -k regex:'conv_tc2_kernel|conv3_halo_kernel' -s 55 -c 12 -o $O/r2_single -f python tools/profile_once.py > $O/r2_ncu_single.log 2>&1
ncu -i $O/r2_single.ncu-rep --page raw --csv > $O/r2_ncu_full_single_raw.csv 2>/dev/null
rm -f $O/r2_single.ncu-rep
ncu --set full --clock-control none -k regex:'stem|mask_decode|nms_kernel|decode_filter8|upsample|sppf' -s 6 -c 7 -o $O/r2_tail -f python tools/profile_once.py > $O/r2_ncu_tail.log 2>&1
ncu -i $O/r2_tail.ncu-rep --page raw --csv > $O/r2_ncu_full_tail_raw.csv 2>/dev/null
rm -f $O/r2_tail.ncu-rep
du -sh $O; ls -la $O/r2_ncu_full_*_raw.csv
