#!/bin/bash
# fused step against context length (standalone harness, graph replays)
out=gpurun_out/${1:-ctx}; mkdir -p $out
for ctx in 96 512 1024 3000; do timeout 60 tools/chainbench step $ctx 30 2>&1 | tee -a $out/time.log | tail -1; done
timeout 60 tools/chainbench chain 1 20 2>&1 | head -1 | tee -a $out/time.log
timeout 60 tools/chainbench chain 8 20 2>&1 | head -1 | tee -a $out/time.log
