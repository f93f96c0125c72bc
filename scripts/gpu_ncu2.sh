#!/bin/bash
out=gpurun_out/ncu2; mkdir -p $out
export CGQ_GEMV_CTAS_PER_SM=4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:w4_gemv -s 4 -c 1 -o $out/prof_lmhead tools/chainbench single 4096 65024 1 1 > $out/ncu.log 2>&1
tail -2 $out/ncu.log
