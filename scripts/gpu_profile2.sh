#!/bin/bash
# round-end ncu evidence for the CURRENT build: launch list of one token step of the bench `value` (113 launches),
# full captures of the decode kernel inside the fused step and of the sampler kernel
tag=${1:-prof}; out=gpurun_out/$tag; mkdir -p $out
B="python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-micro --no-cpu --no-int8"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 400 ncu --metrics $M --clock-control none -k regex:w4_gemv -s 565 -c 113 --csv \
  --log-file $out/launches_token.csv $B > $out/ncu_launch.log 2>&1
tail -1 $out/ncu_launch.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:w4_gemv -s 147 -c 4 \
  -o $out/prof_w4_gemv_fused tools/chainbench step 96 1 > $out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:top_p_sample -s 3 -c 1 \
  -o $out/prof_sampler python scripts/time_sampler.py > $out/ncu_sampler.log 2>&1
ls -la $out
