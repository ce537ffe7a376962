#!/bin/bash
# development aid: bench.py at the given GPU counts, one summary line each
for N in "$@"; do
  if [ "$N" = 1 ]; then CMD="python bench.py"; else
    CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py"; fi
  timeout 400 $CMD --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_n$N.json
  python -c "import json,sys; d=json.load(open('gpurun_out/scale_n$N.json')); print('N', $N, 'ms', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms_per_cycle'],3), 'e2e', round(d['e2e']['ms_per_step'],2), 'value', d['value'])"
done
