#!/bin/bash
# launch list (cold-cache, serialised) of one batched beam-decode step; usage: ncu_beam.sh [k ...]
cd /root/repo
mkdir -p gpurun_out
for k in ${@:-5}; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'gemv_beams|beam_attention|norm_beams' --launch-skip 452 --launch-count 226 --csv \
    --log-file gpurun_out/r2_beam_launches_k$k.csv python scripts/beam_step_once.py $k > gpurun_out/beam_ncu_k$k.log 2>&1
  tail -2 gpurun_out/beam_ncu_k$k.log
done
