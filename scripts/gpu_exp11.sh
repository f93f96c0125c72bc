#!/bin/bash
# M=1 consumer restructure (stages as MMA columns): parity + timings
out=gpurun_out/exp11; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_fused_decode.py -x -q -m gpu > $out/pytest_fused.log 2>&1; echo "rc=$?" >> $out/pytest_fused.log; tail -5 $out/pytest_fused.log
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "int4 and not tcgen05 and not umma" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -5 $out/pytest.log
{
echo "== chain"; timeout 60 tools/chainbench chain 1 20
for s in "4096 65024" "4096 27392" "4096 13696" "4096 4608" "4096 4096" "13696 4096"; do timeout 60 tools/chainbench single $s 1 10; done
echo "== fused step"; timeout 60 tools/chainbench step 96 30
} > $out/log.txt 2>&1
cat $out/log.txt
