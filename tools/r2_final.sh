#!/bin/bash
# Round-end validation on one B200: GPU parity tests, smoke(), the default bench line and the reference arm.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2f_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r2f_pytest.log
tail -3 $O/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2f_smoke.log 2>&1; tail -1 $O/r2f_smoke.log
python bench.py > $O/r2f_bench_default.log 2> $O/r2f_bench_default.err; echo "bench exit $?"
tail -1 $O/r2f_bench_default.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pinned', round(d['e2e']['pinned_frames']['value']), 'frac', round(d['roofline']['frac'], 3), 'traffic', d['roofline']['traffic'], d['e2e']['call_ms_min_median_max']['pageable']['repetitions'], 'b1', d['p50_frame_latency_ms_b1'], d['p50_predict_call_ms_b1'], d['clocks'], d['cpu_baseline'], d['gpu_launches'])"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2f_bench_ref.log 2>&1; tail -1 $O/r2f_bench_ref.log | cut -c1-300
