#!/bin/bash
# ncu evidence for profiles/: per-launch time + DRAM bytes of one token step, full captures of the
# decode and prefill kernels.  Usage (under gpurun): bash scripts/gpu_profile.sh <tag>
tag=${1:-prof}; out=gpurun_out/$tag; mkdir -p $out
B="python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-micro --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:w4_gemv -s 565 -c 113 --csv --log-file $out/launches_token.csv $B > $out/ncu_launch.log 2>&1
tail -2 $out/ncu_launch.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:w4_gemv -s 565 -c 4 \
  -o $out/prof_w4_gemv $B > $out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wq_gemm_tc -s 4 -c 1 \
  -o $out/prof_tc tools/chainbench single 4096 27392 2048 1 > $out/ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wq_gemm_tc -s 4 -c 1 \
  -o $out/prof_tc_m128 tools/chainbench single 4096 27392 128 1 > $out/ncu_tc128.log 2>&1
ls -la $out
