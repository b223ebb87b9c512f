#!/bin/bash
# Full GPU suite + the default workload's line (value, e2e, B=1 latencies) + small-batch points.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2q_pytest.log 2>&1; tail -3 $O/r2q_pytest.log
for w in yolov8s-seg-640-b64 yolov8s-seg-640-b8 yolov8n-seg-640-b1 "$@"; do
  python bench.py --steps 100 --no-cpu-baseline --workload $w > $O/r2q_$w.log 2> $O/r2q_$w.err || { tail -5 $O/r2q_$w.err; continue; }
  tail -1 $O/r2q_$w.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['config']['workload'], round(d['value']), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), round(d['e2e']['pinned_frames']['value']), 'b1 dev/call ms', round(d['p50_frame_latency_ms_b1'], 3), round(d['p50_predict_call_ms_b1'], 3), 'frac', round(d['roofline']['frac'], 3))"
done
