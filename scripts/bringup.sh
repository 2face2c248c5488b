#!/bin/bash
# Runs every bring-up group in its own process (a trap poisons the CUDA context) under a timeout.
# $1 = "dbg" to use the bounded-spin debug library.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/bringup_${1:-rel}.log
: > $LOG
if [ "$1" = "dbg" ]; then export MERLIN_B200_LIB=$PWD/merlin_b200/libmerlin_b200_dbg.so; fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
shift
for g in "$@"; do
  timeout 240 python scripts/bringup.py $g >> $LOG 2>&1
  echo "--- exit code $g: $?" >> $LOG
done
tail -c 6000 $LOG
