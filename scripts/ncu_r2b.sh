#!/bin/bash
# Round-2 (second half) `--set full` captures of the kernels that changed: run under gpurun, ONE GPU.
set -x
O=gpurun_out
NCU="ncu --clock-control none --set full --import-source on"
$NCU -k regex:mlp_pair -s 6 -c 2 -o $O/r2b_full_mlp_pair_c2 -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:mlp_pair -s 2 -c 2 -o $O/r2b_full_mlp_pair_c3 -f python scripts/_c3_step.py > /dev/null 2>&1
$NCU -k regex:lbs_warp_kernel -s 2 -c 1 -o $O/r2b_full_lbs_warp -f python scripts/_c3_step.py > /dev/null 2>&1
$NCU -k regex:gemm_pair_kernel -s 1 -c 1 -o $O/r2b_full_gemm_pair_1k -f python scripts/time_wide.py 524288 > /dev/null 2>&1
for m in fwd1k dgrad1k wgrad1k; do
  $NCU -k regex:'gemm_tma|wgrad_tma' -s 1 -c 1 -o $O/r2b_full_gemm_$m -f python scripts/ncu_gemm.py $m > /dev/null 2>&1
done
$NCU -k regex:colsum_f16_v8 -s 1 -c 1 -o $O/r2b_full_colsum -f python scripts/ncu_gemm.py colsum > /dev/null 2>&1
$NCU -k regex:resample_level -s 5 -c 1 -o $O/r2b_full_resample_l1 -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
ls -la $O/r2b_*.ncu-rep
