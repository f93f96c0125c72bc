#!/bin/bash
out=gpurun_out/exp3; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "int4 or module or edge" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -6 $out/pytest.log
{
echo "== chain default"; timeout 120 tools/chainbench chain 1 20
for s in "4096 4096" "4096 4608" "4096 27392" "13696 4096" "4096 65024" "4096 37888"; do timeout 120 tools/chainbench single $s 1 10; done
echo "== chain M=8"; timeout 120 tools/chainbench chain 8 10
echo "== trace"; timeout 120 tools/chainbench trace 1 | head -36
for st in 3 5 6 8; do echo "== STAGES=$st"; CGQ_GEMV_STAGES=$st timeout 120 tools/chainbench chain 1 20 | head -1; done
for z in 4; do echo "== Z=$z"; CGQ_GEMV_Z=$z timeout 120 tools/chainbench chain 1 20 | head -1; done
} > $out/log.txt 2>&1
cat $out/log.txt
