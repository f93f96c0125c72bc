#!/bin/bash
# ncu evidence for profiles/: per-launch time + DRAM bytes of one token step (113-linear chain through bench.py
# and the 142-launch fused step through tools/chainbench), full captures of the decode / attention kernels.
# Usage (under gpurun): bash scripts/gpu_profile.sh <tag>
tag=${1:-prof}; out=gpurun_out/$tag; mkdir -p $out
B="python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-micro --no-cpu --no-int8"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 ncu --metrics $M --clock-control none -k regex:w4_gemv -s 565 -c 113 --csv \
  --log-file $out/launches_token.csv $B > $out/ncu_launch.log 2>&1
tail -1 $out/ncu_launch.log | cut -c1-200
timeout 600 ncu --metrics $M --clock-control none -s 142 -c 142 --csv \
  --log-file $out/launches_fused_step.csv tools/chainbench step 96 1 > $out/ncu_fused.log 2>&1
tail -1 $out/ncu_fused.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:w4_gemv -s 147 -c 4 \
  -o $out/prof_w4_gemv_fused tools/chainbench step 96 1 > $out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_attn -s 29 -c 1 \
  -o $out/prof_attn tools/chainbench step 96 1 > $out/ncu_attn.log 2>&1
ls -la $out
