#!/bin/bash
# Round-2 profiler captures (run under gpurun, ONE GPU): launch lists of the three step kinds + one `--set full` capture per kernel.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
# launch lists (per-launch durations; cold-cache, serialised: shares, not absolutes)
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r2_launches_c2_fp16.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/r2_launches_fp16x3.csv python scripts/time_modes.py fp16x3 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 2500 --csv --log-file $O/r2_launches_train.csv python scripts/time_train.py > /dev/null 2>&1
# full captures, one launch each
for m in fwd dgrad split wgrad; do
  $NCU --set full --import-source on -k regex:'gemm_tma|wgrad_tma' -s 1 -c 1 -o $O/r2_full_gemm_$m -f python scripts/ncu_gemm.py $m > /dev/null 2>&1
done
$NCU --set full --import-source on -k regex:mlp_pair -s 6 -c 2 -o $O/r2_full_mlp_pair -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
$NCU --set full --import-source on -k regex:ipe_features_rm16 -s 2 -c 1 -o $O/r2_full_ipe_rm16 -f python scripts/time_modes.py fp16x3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:resample_level -s 4 -c 1 -o $O/r2_full_resample -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
$NCU --set full --import-source on -k regex:composite_mip360_kernel -s 4 -c 1 -o $O/r2_full_composite -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
$NCU --set full --import-source on -k regex:'lbs_warp_kernel' -s 2 -c 1 -o $O/r2_full_lbs_warp -f python scripts/_c3_step.py > /dev/null 2>&1
$NCU --set full --import-source on -k regex:'lbs_backward' -s 0 -c 2 -o $O/r2_full_lbs_bwd -f python -m pytest tests/test_gpu_models.py -q -m gpu -k human_s2_training > /dev/null 2>&1
ls -la $O/*.ncu-rep $O/r2_launches_*.csv
