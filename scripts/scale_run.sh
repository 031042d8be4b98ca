#!/bin/bash
# usage: scale_run.sh N workload...   -> gpurun_out/scale_<workload>_n<N>.json (one bench line each)
N=$1; shift
for w in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --workload $w --no-e2e --no-cpu-baseline > gpurun_out/scale_${w}_n${N}.log 2>&1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --workload $w > gpurun_out/scale_${w}_n${N}.log 2>&1
  fi
  echo "$w n=$N rc=$?"
  tail -1 gpurun_out/scale_${w}_n${N}.log | python scripts/bench_line.py
  tail -1 gpurun_out/scale_${w}_n${N}.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    g = d.get('gathered')
    print('   gathered:', g and round(g['ms_per_step'], 2), 'ms;', d['config']['sharding'][:150])
except Exception as e:
    print('   (no json)', e)
"
done
