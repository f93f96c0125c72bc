#!/bin/bash
out=gpurun_out/exp8; mkdir -p $out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "int4 and not tcgen05" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -6 $out/pytest.log
{
echo "== default (auto wide)"; timeout 60 tools/chainbench chain 1 20
for s in "4096 65024" "4096 27392" "4096 13696" "4096 4096" "13696 4096"; do timeout 60 tools/chainbench single $s 1 10; done
echo "== wide off"; CGQ_GEMV_WIDE=0 timeout 60 tools/chainbench chain 1 20 | head -1
for s in "4096 65024" "4096 27392" "4096 13696"; do CGQ_GEMV_WIDE=0 timeout 60 tools/chainbench single $s 1 10; done
echo "== wide on everywhere"; CGQ_GEMV_WIDE=1 timeout 60 tools/chainbench chain 1 20 | head -1
for s in "4096 13696" "4096 4096" "13696 4096"; do CGQ_GEMV_WIDE=1 timeout 60 tools/chainbench single $s 1 10; done
for st in 3 6 8; do echo "== STAGES=$st"; CGQ_GEMV_STAGES=$st timeout 60 tools/chainbench chain 1 20 | head -1; CGQ_GEMV_STAGES=$st timeout 60 tools/chainbench single 4096 65024 1 10; done
} > $out/log.txt 2>&1
cat $out/log.txt
