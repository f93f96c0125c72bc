#!/bin/bash
out=gpurun_out/ncu3; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:w4_gemv_umma -s 4 -c 1 -o $out/prof_umma tools/chainbench single 4096 65024 1 1 > $out/ncu.log 2>&1
tail -2 $out/ncu.log
