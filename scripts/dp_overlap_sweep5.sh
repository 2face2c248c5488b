#!/bin/bash
# N=4: the N=8 default (registered arena, 8 CTAs) exposes 51 ms here; which setting hides the exchange at 4 ranks?
N=${1:-4}
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 scripts/train_timeline.py --out gpurun_out/r2_tl_n${N}_$tag > gpurun_out/r2_tl_n${N}_$tag.log 2>&1
  tail -1 gpurun_out/r2_tl_n${N}_$tag.log | cut -c1-330
  rm -f gpurun_out/r2_tl_n${N}_$tag.trace.json.gz gpurun_out/r2_tl_n${N}_${tag}_kernels.csv.gz
}
run reg_c16_r16  MERLIN_DP_REGISTER=1 MERLIN_DP_CTAS=16 MERLIN_DP_SM_RESERVE=16 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING,REG
grep -i "NCCL INFO" gpurun_out/r2_tl_n${N}_reg_c16_r16.log | grep "\[0\]" | grep -i "regist\|AllReduce:\|nvls" | sort | uniq -c | sort -rn | head -8 | cut -c1-200
run c16_r16      MERLIN_DP_REGISTER=0 MERLIN_DP_CTAS=16 MERLIN_DP_SM_RESERVE=16
run reg_c24_r24  MERLIN_DP_REGISTER=1 MERLIN_DP_CTAS=24 MERLIN_DP_SM_RESERVE=24
run torch        MERLIN_DP_BACKEND=torch
