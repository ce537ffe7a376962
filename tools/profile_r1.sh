#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run, and one --set full capture each of
# the descend and EMA kernels (B200_PROFILING.md recipe).  Numbers printed by bench.py under ncu are
# not bench values.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1d_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:descend_lockstep -s 1 -c 1 -f -o gpurun_out/r1d_descend \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1d_ncu_descend.log 2>&1
ncu --set full --clock-control none -k regex:ema_kernel -s 1 -c 1 -f -o gpurun_out/r1d_ema \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1d_ncu_ema.log 2>&1
ls -la gpurun_out/*.ncu-rep
