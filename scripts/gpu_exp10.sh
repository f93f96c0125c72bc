#!/bin/bash
# fused decode step + L2 prefetch hints: parity tests, timings with hints on/off and prefetch caps
out=gpurun_out/exp10; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_fused_decode.py -x -q -m gpu > $out/pytest_fused.log 2>&1; echo "rc=$?" >> $out/pytest_fused.log; tail -15 $out/pytest_fused.log
{
for h in 0 1; do
echo "== HINTS=$h chain"; CGQ_BENCH_HINTS=$h timeout 60 tools/chainbench chain 1 20
for s in "4096 65024" "4096 27392" "4096 13696" "4096 4608" "4096 4096" "13696 4096"; do CGQ_BENCH_HINTS=$h timeout 60 tools/chainbench single $s 1 10; done
echo "== HINTS=$h fused step"; CGQ_BENCH_HINTS=$h timeout 60 tools/chainbench step 96 30
done
for mb in 16 32 64 96; do echo "== PF_MB=$mb"; CGQ_PF_MB=$mb timeout 60 tools/chainbench chain 1 20 | head -1; CGQ_PF_MB=$mb timeout 60 tools/chainbench step 96 30; done
echo "== steptrace (hints on)"; timeout 60 tools/chainbench steptrace 96 | head -56
} > $out/log.txt 2>&1
cat $out/log.txt
