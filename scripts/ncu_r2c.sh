#!/bin/bash
# Launch lists of the final kernels (per-launch durations; cold-cache, serialised: shares, not absolutes).  Run under gpurun, ONE GPU.
O=gpurun_out
NCU="ncu --clock-control none --metrics gpu__time_duration.sum --csv"
$NCU -c 400 --log-file $O/r2c_launches_c2_fp16.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
$NCU -c 2500 --log-file $O/r2c_launches_train.csv python scripts/time_train.py > /dev/null 2>&1
$NCU -c 200 --log-file $O/r2c_launches_c3.csv python scripts/_c3_step.py > /dev/null 2>&1
ls -la $O/r2c_launches_*.csv
