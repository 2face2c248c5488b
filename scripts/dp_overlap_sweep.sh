#!/bin/bash
# N-GPU sweep of the gradient-exchange settings on bench.py's training step (scripts/train_timeline.py):
#   usage: scripts/dp_overlap_sweep.sh [N=2] ; writes gpurun_out/r2_tl_<tag>.{json,_kernels.csv,trace.json.gz}
N=${1:-2}
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 scripts/train_timeline.py --out gpurun_out/r2_tl_n${N}_$tag > gpurun_out/r2_tl_n${N}_$tag.log 2>&1
  tail -1 gpurun_out/r2_tl_n${N}_$tag.log | cut -c1-700
}
run torch      MERLIN_DP_BACKEND=torch
run lib_c8_r8  MERLIN_DP_CTAS=8  MERLIN_DP_SM_RESERVE=8
run lib_c8_r16 MERLIN_DP_CTAS=8  MERLIN_DP_SM_RESERVE=16
run lib_c4_r8  MERLIN_DP_CTAS=4  MERLIN_DP_SM_RESERVE=8
run lib_c16_r16 MERLIN_DP_CTAS=16 MERLIN_DP_SM_RESERVE=16
run lib_c0_r0  MERLIN_DP_CTAS=0  MERLIN_DP_SM_RESERVE=0
rm -f gpurun_out/r2_tl_n${N}_lib_c4_r8.trace.json.gz gpurun_out/r2_tl_n${N}_lib_c16_r16.trace.json.gz gpurun_out/r2_tl_n${N}_lib_c0_r0.trace.json.gz gpurun_out/r2_tl_n${N}_lib_c8_r16.trace.json.gz
