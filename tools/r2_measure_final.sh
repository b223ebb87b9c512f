#!/bin/bash
# Final round-2 bench lines and per-op tables of every workload on ONE B200 (no profiler): profiles/r2_bench_lines.jsonl,
# profiles/r2_ops_*.csv are copies of what this writes under gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2_bench_lines.jsonl
python bench.py --dump-ops $O/r2_ops_yolov8s-seg-640-b64.csv > $O/r2_bench_default.log 2> $O/r2_bench_default.err
tail -1 $O/r2_bench_default.log >> $O/r2_bench_lines.jsonl
for w in yolov10n-640-b32 yolov8n-seg-640-b64 yolov8m-seg-1080p-b16 yolov8x-seg-640-b32 yolo11n-seg-640-b64 yolo11s-seg-640-b64 yolo11x-seg-640-b32 yolov8s-seg-640-b1; do
  python bench.py --workload $w --steps 100 --warmup 3 --no-cpu-baseline --dump-ops $O/r2_ops_$w.csv > $O/r2_bench_$w.log 2>/dev/null
  tail -1 $O/r2_bench_$w.log >> $O/r2_bench_lines.jsonl
done
python - <<PY
import json
for l in open("$O/r2_bench_lines.jsonl"):
    d = json.loads(l)
    r = d["roofline"]
    print(d["config"]["workload"], round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["pinned_frames"]["value"]),
          "frac", round(r["frac"], 3), round(r["frac_of_sustained_peak"], 3), "2c", round(r["two_ceiling"]["frac"], 3), "conv TF", round(r["achieved"]), "b1", round(d["p50_frame_latency_ms_b1"], 3),
          round(d["p50_predict_call_ms_b1"], 3), d["clocks"]["sm_mhz"], "det", d["config"]["detections_per_step"], {k: round(v, 3) for k, v in r["step_breakdown_ms"].items()},
          "handoff", (d["e2e"].get("index_mask_handoff") or {}).get("ms_per_step"))
PY
