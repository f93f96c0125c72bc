#!/bin/bash
out=gpurun_out/tc1; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tcgen05 or prefill" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -15 $out/pytest.log
{
for s in "4096 4608 128" "4096 13696 128" "4096 27392 128" "4096 4608 2048" "4096 13696 2048" "4096 27392 2048" "13696 4096 2048" "4096 65024 2048"; do timeout 120 tools/chainbench single $s 5; done
} > $out/perf.log 2>&1
cat $out/perf.log
