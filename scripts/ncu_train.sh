#!/bin/bash
# ncu captures of the training kernels (one launch each) + the launch list of one training step.
set -x
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:"attn_bwd_kernel|rmsnorm_bwd_warp_kernel|cross_entropy_bwd_kernel" -c 5 -f -o gpurun_out/r2_train_bwd python scripts/train_one_step.py > gpurun_out/r2_ncu_train_bwd.log 2>&1
$NCU -k regex:"adamw_kernel|sumsq_kernel|attn_delta_kernel" -c 3 -f -o gpurun_out/r2_train_opt python scripts/train_one_step.py > gpurun_out/r2_ncu_train_opt.log 2>&1
$NCU --kernel-name-base demangled -k regex:"gemm_bf16_kernel<\(int\)2, \(bool\)1, \(bool\)1>" -c 1 -f -o gpurun_out/r2_train_wgrad python scripts/train_one_step.py > gpurun_out/r2_ncu_train_wgrad.log 2>&1
$NCU --kernel-name-base demangled -k regex:"gemm_bf16_kernel<\(int\)2, \(bool\)0, \(bool\)1>" -c 1 -f -o gpurun_out/r2_train_dgrad python scripts/train_one_step.py > gpurun_out/r2_ncu_train_dgrad.log 2>&1
S_LIST=4096 $NCU -k regex:"attn_fwd2_kernel" -c 1 --launch-skip 3 -f -o gpurun_out/r2_attn_fwd2 python scripts/attn_fwd_bench.py > gpurun_out/r2_ncu_attn_fwd2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_train_launches.csv python scripts/train_one_step.py > gpurun_out/r2_ncu_train_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep
