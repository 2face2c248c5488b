#!/bin/bash
# Second sweep: NCCL protocol / CTA budget (see dp_overlap_sweep.sh).  usage: scripts/dp_overlap_sweep2.sh [N=2]
N=${1:-2}
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 scripts/train_timeline.py --out gpurun_out/r2_tl_n${N}_$tag > gpurun_out/r2_tl_n${N}_$tag.log 2>&1
  tail -1 gpurun_out/r2_tl_n${N}_$tag.log | cut -c1-420
  rm -f gpurun_out/r2_tl_n${N}_$tag.trace.json.gz
}
nvidia-smi topo -m > gpurun_out/r2_topo_n${N}.txt 2>&1
run simple_c8_r8   NCCL_PROTO=Simple MERLIN_DP_CTAS=8 MERLIN_DP_SM_RESERVE=8 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH,TUNING,ENV
grep -i "NCCL INFO" gpurun_out/r2_tl_n${N}_simple_c8_r8.log | grep -i "nvls\|channel\|algo\|proto\|P2P\|NET\|Connected\|comm 0x\|nRanks\|CTA\|Ring\|Tree" | head -60 > gpurun_out/r2_nccl_info_n${N}.txt
run simple_c4_r4   NCCL_PROTO=Simple MERLIN_DP_CTAS=4 MERLIN_DP_SM_RESERVE=4
run simple_c2_r2   NCCL_PROTO=Simple MERLIN_DP_CTAS=2 MERLIN_DP_SM_RESERVE=2
run simple_c8_r4   NCCL_PROTO=Simple MERLIN_DP_CTAS=8 MERLIN_DP_SM_RESERVE=4
run simple_c16_r16 NCCL_PROTO=Simple MERLIN_DP_CTAS=16 MERLIN_DP_SM_RESERVE=16
run simple_c0_r0   NCCL_PROTO=Simple MERLIN_DP_CTAS=0 MERLIN_DP_SM_RESERVE=0
