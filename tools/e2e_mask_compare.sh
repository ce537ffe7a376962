#!/bin/bash
# development aid: e2e step time with the full-record download vs the 16-byte field download
for m in all hdm; do
  timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-mask $m 2>/dev/null | tail -1 > gpurun_out/e2e_$m.json
  python -c "import json; d=json.load(open('gpurun_out/e2e_$m.json')); print('$m', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_device_consumers']['ms_per_step'])"
done
