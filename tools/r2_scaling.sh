#!/bin/bash
# Multi-GPU points of BASELINE.json's configs on ONE box (torchrun, one rank per GPU): the default workload (C3), C4
# (yolov8m-seg, 1080p frames letterboxed to 736x1280, 16 frames per GPU) and C5 (yolov8x-seg, 32 frames per GPU = 256 per
# step on 8 GPUs, index-mask hand-off + peer push to rank 0 inside the timed e2e region).  usage: r2_scaling.sh N [workload ...]
N=$1
shift
WL="${@:-yolov8s-seg-640-b64 yolov8m-seg-1080p-b16 yolov8x-seg-640-b32}"
cd "$(dirname "$0")/.."
O=gpurun_out
P=29600
for w in $WL; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N \
      --steps 60 --warmup 3 --workload $w > $O/r2_scale_${w}_n$N.log 2> $O/r2_scale_${w}_n$N.err
  tail -1 $O/r2_scale_${w}_n$N.log >> $O/r2_scaling_lines.jsonl
  tail -1 $O/r2_scale_${w}_n$N.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['config']['workload'], 'n_gpus', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'pinned', round(d['e2e']['pinned_frames']['value']), d['e2e'].get('index_mask_handoff', {}) and d['e2e']['index_mask_handoff'].get('peer_push_to_rank0_mailbox'))"
done
