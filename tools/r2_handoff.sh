#!/bin/bash
# Hand-off tests + the two workloads whose e2e / selection work changed (C5 with the hand-off inside, C4 with detections).
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_handoff.py tests/test_sharded_predictor.py "tests/test_gpu_engine.py::test_config_c4_yolov8m_seg_1080p_b2_strict" -m gpu -x -q > $O/r2h_pytest.log 2>&1
tail -4 $O/r2h_pytest.log
for w in yolov8x-seg-640-b32 yolov8m-seg-1080p-b16; do
  python bench.py --steps 60 --no-cpu-baseline --workload $w > $O/r2h_$w.log 2> $O/r2h_$w.err || { tail -5 $O/r2h_$w.err; continue; }
  tail -1 $O/r2h_$w.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['config']['workload'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), round(d['e2e']['pinned_frames']['value']), d['e2e']['index_mask_handoff'], d['config']['detections_per_step'], d['roofline']['step_breakdown_ms'], d['e2e']['call_ms_min_median_max'])"
done
