#!/bin/bash
out=gpurun_out/ncu1; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:w4_gemv -s 12 -c 2 -o $out/prof_single tools/chainbench single 4096 37888 1 1 > $out/ncu.log 2>&1
tail -3 $out/ncu.log
