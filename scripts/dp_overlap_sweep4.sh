#!/bin/bash
# N=8: does NCCL user-buffer registration of the gradient arena (NVLS in-place) let fewer CTAs keep up?
N=${1:-8}
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 scripts/train_timeline.py --out gpurun_out/r2_tl_n${N}_$tag > gpurun_out/r2_tl_n${N}_$tag.log 2>&1
  tail -1 gpurun_out/r2_tl_n${N}_$tag.log | cut -c1-330
  rm -f gpurun_out/r2_tl_n${N}_$tag.trace.json.gz
}
run reg_c8_r8    MERLIN_DP_REGISTER=1 MERLIN_DP_CTAS=8  MERLIN_DP_SM_RESERVE=8 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING,REG
grep -i "NCCL INFO" gpurun_out/r2_tl_n${N}_reg_c8_r8.log | grep "\[0\]" | grep -i "regist\|AllReduce:\|nvls" | sort | uniq -c | sort -rn | head -12 | cut -c1-220
run reg_c12_r12  MERLIN_DP_REGISTER=1 MERLIN_DP_CTAS=12 MERLIN_DP_SM_RESERVE=12
run reg_c16_r16  MERLIN_DP_REGISTER=1 MERLIN_DP_CTAS=16 MERLIN_DP_SM_RESERVE=16
run c24_r24      MERLIN_DP_CTAS=24 MERLIN_DP_SM_RESERVE=24
run c16_r16      MERLIN_DP_CTAS=16 MERLIN_DP_SM_RESERVE=16
