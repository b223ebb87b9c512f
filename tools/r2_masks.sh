#!/bin/bash
# Mask-decode change check: the mask parity tests, then the step breakdown of the s-seg, x-seg and m-seg workloads.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_gpu_engine.py tests/test_gpu_v11.py tests/test_handoff.py -m gpu -x -q > $O/r2m_pytest.log 2>&1
tail -3 $O/r2m_pytest.log
grep -h mask_mismatch_px $O/parity_report.jsonl | python -c "
import sys, json
rows = [json.loads(l) for l in sys.stdin]
print('mask rows', len(rows), 'total mismatching px', sum(r['mask_mismatch_px'] for r in rows), 'max', max(r['mask_mismatch_px'] for r in rows))"
for w in yolov8s-seg-640-b64 yolov8x-seg-640-b32 yolov8m-seg-1080p-b16; do
  python bench.py --steps 60 --no-cpu-baseline --workload $w > $O/r2m_$w.log 2> $O/r2m_$w.err || { tail -5 $O/r2m_$w.err; continue; }
  tail -1 $O/r2m_$w.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['config']['workload'], round(d['value']), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), round(d['e2e']['pinned_frames']['value']), d['roofline']['step_breakdown_ms'])"
done
