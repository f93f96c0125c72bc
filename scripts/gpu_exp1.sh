#!/bin/bash
# experiment: decode-chain timing under ring-depth / CTA-per-SM / PDL variants + in-kernel timeline
out=gpurun_out/exp1; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "int4" > $out/pytest.log 2>&1; tail -3 $out/pytest.log
{
for cfg in "5 2 1" "8 2 1" "4 2 1" "3 2 1" "6 1 1" "10 1 1" "5 2 0" "3 3 1" "4 4 1"; do
  set -- $cfg
  echo "== STAGES=$1 CTAS_PER_SM=$2 PDL=$3"
  CGQ_GEMV_STAGES=$1 CGQ_GEMV_CTAS_PER_SM=$2 CGQ_PDL=$3 timeout 120 tools/chainbench chain 1 20
done
echo "== singles (default cfg)"
for s in "4096 4096" "4096 4608" "4096 27392" "13696 4096" "4096 65024"; do timeout 120 tools/chainbench single $s 1 10; done
echo "== trace (default cfg)"
timeout 120 tools/chainbench trace 1
echo "== trace STAGES=8 (no co-residency)"
CGQ_GEMV_STAGES=8 timeout 120 tools/chainbench trace 1
} > $out/chain.log 2>&1
cat $out/chain.log
