#!/bin/bash
# End-of-round measurement pass on one B200 (through gpurun): bench lines of every workload, per-op tables, the ncu
# launch list of the bench command and the per-launch DRAM traffic of one forward pass.
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py --dump-ops $O/ops_final_yolov8s-seg-640-b64.csv > $O/bench_final_default.log 2>&1
for w in yolov10n-640-b32 yolov8n-seg-640-b64 yolov8m-seg-1080p-b16 yolov8x-seg-640-b32 yolo11n-seg-640-b64 yolo11s-seg-640-b64 yolo11x-seg-640-b32; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_final_$w.csv > $O/bench_final_$w.log 2>&1
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r1e_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
bash tools/ncu_traffic.sh
for f in $O/bench_final_*.log; do tail -1 $f | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['pageable_frames']['value'],1), 'frac', round(d['roofline']['frac'],3), 'b1', round(d['p50_frame_latency_ms_b1'],3))"; done
# full-set capture of the conv launches around the prototype branch (tcgen05 / TMA evidence for the current kernels)
python tools/profile_once.py > $O/ncu_once_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'conv' -s 58 -c 10 -o $O/r1e_conv_tail -f python tools/profile_once.py > $O/ncu_once_conv.log 2>&1
ncu -i $O/r1e_conv_tail.ncu-rep --page raw --csv > $O/r1e_ncu_full_conv_tail_raw.csv 2>/dev/null
rm -f $O/r1e_conv_tail.ncu-rep
python tools/nms_probe.py > $O/r1e_nms_probe.log 2>&1
du -sh $O
