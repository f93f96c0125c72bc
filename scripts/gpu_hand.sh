#!/bin/bash
# experimental tile-granular hand-over: which polling variant / which hand-overs cost what
out=gpurun_out/${1:-hand}; mkdir -p $out
for cfg in "0 20 7" "1 20 7" "3 20 7" "3 200 7" "3 20 1" "3 20 2" "3 20 4"; do
  set -- $cfg
  CGQ_HAND_MODE=$1 CGQ_HAND_SLEEP=$2 CGQ_HAND_MASK=$3 timeout 120 python scripts/time_fused_step.py 2>&1 | grep "fused step" | tee -a $out/time.log
done
