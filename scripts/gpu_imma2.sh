#!/bin/bash
# timeline + tunables of the IMMA decode kernel
tag=${1:-imma2}; out=gpurun_out/$tag; mkdir -p $out
{ timeout 60 tools/chainbench trace 1 | head -75
for st in 3 4 5 6 8; do echo "== CGQ_GEMV_STAGES=$st"; CGQ_GEMV_STAGES=$st timeout 60 tools/chainbench chain 1 20 | head -1; done
for c in 3 4; do echo "== CGQ_GEMV_CTAS_PER_SM=$c"; CGQ_GEMV_CTAS_PER_SM=$c timeout 60 tools/chainbench chain 1 20 | head -1; done
for z in 1 2 4 8; do echo "== CGQ_GEMV_Z=$z"; for s in "4096 4608" "4096 4096" "13696 4096" "4096 27392" "4096 65024"; do CGQ_GEMV_Z=$z timeout 60 tools/chainbench single $s 1 10; done; done
} 2>&1 | tee $out/log.txt
