#!/bin/bash
out=gpurun_out/exp2; mkdir -p $out
{
echo "== chain default"; timeout 120 tools/chainbench chain 1 20
echo "== singles: lockstep hypothesis (N = 128*296, 128*148) vs ragged"
for s in "4096 37888" "4096 18944" "4096 65024" "4096 27392" "4096 4096" "13696 4096"; do timeout 120 tools/chainbench single $s 1 10; done
echo "== trace"; timeout 120 tools/chainbench trace 1 | head -40
} > $out/log.txt 2>&1
cat $out/log.txt
