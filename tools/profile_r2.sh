#!/bin/bash
# round-2 ncu evidence for profiles/ (B200_PROFILING.md recipe).  Numbers printed by a run under ncu are never bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
# launch list of a short bench run (same command as the bench, fewer steps)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traffic > $O/r2_bench_under_ncu.log 2>&1
# one --set full capture per descend kernel in the configuration that uses it
ncu --set full --clock-control none --import-source on -k regex:descend_lockstep -s 2 -c 1 -f -o $O/r2_dense_8192 \
    python tools/tune_descend.py 16 0:0:0 > $O/r2_ncu_dense.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:descend_lockstep -s 2 -c 1 -f -o $O/r2_mid_4096 \
    python tools/tune_descend.py 8 0:0:0 > $O/r2_ncu_mid4096.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:descend_lockstep -s 2 -c 1 -f -o $O/r2_small_2048 \
    python tools/tune_descend.py 4 0:0:0 > $O/r2_ncu_small2048.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:descend_group -s 2 -c 1 -f -o $O/r2_group_512 \
    python tools/tune_descend.py 1 0:0:0 > $O/r2_ncu_group512.log 2>&1
ncu --set full --clock-control none -k regex:ema_kernel -s 2 -c 1 -f -o $O/r2_ema \
    python tools/tune_descend.py 16 0:0:0 > $O/r2_ncu_ema.log 2>&1
ls -la $O/r2_*.ncu-rep
