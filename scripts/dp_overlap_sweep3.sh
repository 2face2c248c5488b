#!/bin/bash
# N-GPU validation of the gradient exchange settings (see dp_overlap_sweep.sh).  usage: scripts/dp_overlap_sweep3.sh [N=8]
N=${1:-8}
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 scripts/train_timeline.py --out gpurun_out/r2_tl_n${N}_$tag > gpurun_out/r2_tl_n${N}_$tag.log 2>&1
  tail -1 gpurun_out/r2_tl_n${N}_$tag.log | cut -c1-480
}
nvidia-smi topo -m > gpurun_out/r2_topo_n${N}.txt 2>&1
run lib_c8_r8   MERLIN_DP_CTAS=8 MERLIN_DP_SM_RESERVE=8 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING
grep -i "NCCL INFO" gpurun_out/r2_tl_n${N}_lib_c8_r8.log | grep -i "nvls\|nChannels\|algo\|proto\|Connected\|comm 0x" | grep "\[0\]" | head -40 > gpurun_out/r2_nccl_info_n${N}.txt
run torch       MERLIN_DP_BACKEND=torch
run lib_c16_r16 MERLIN_DP_CTAS=16 MERLIN_DP_SM_RESERVE=16
run lib_c4_r4   MERLIN_DP_CTAS=4 MERLIN_DP_SM_RESERVE=4
rm -f gpurun_out/r2_tl_n${N}_lib_c16_r16.trace.json.gz gpurun_out/r2_tl_n${N}_lib_c4_r4.trace.json.gz
grep -c "ncclDevKernel" gpurun_out/r2_tl_n${N}_lib_c8_r8_kernels.csv; grep "ncclDevKernel" gpurun_out/r2_tl_n${N}_lib_c8_r8_kernels.csv | head -3 | cut -c1-160
