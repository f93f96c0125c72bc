#!/bin/bash
out=gpurun_out/exp4; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "int4 or module or edge" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -12 $out/pytest.log
{
echo "== chain (m1 kernel)"; timeout 120 tools/chainbench chain 1 20
for s in "4096 4096" "4096 4608" "4096 27392" "13696 4096" "4096 65024" "4096 37888"; do timeout 120 tools/chainbench single $s 1 10; done
echo "== chain (m1 disabled)"; CGQ_GEMV_M1=0 timeout 120 tools/chainbench chain 1 20 | head -1
echo "== trace"; timeout 120 tools/chainbench trace 1 | head -30
for st in 3 5 6 8; do echo "== STAGES=$st"; CGQ_GEMV_STAGES=$st timeout 120 tools/chainbench chain 1 20 | head -1; done
} > $out/log.txt 2>&1
cat $out/log.txt
