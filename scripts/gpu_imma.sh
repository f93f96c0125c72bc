#!/bin/bash
# integer-MMA decode arithmetic (gemv_w4.cu kImma): parity first, then A/B timing against the subnormal-operand path
tag=${1:-imma}; out=gpurun_out/$tag; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_decode.py -x -q -m gpu -k "int4 or stress or fused or gemv or program or tp_shard or decode or grad" 2>&1 | tail -15 | tee $out/pytest.txt
for ar in 0 2; do
  echo "== CGQ_GEMV_ARITH=$ar (0 = imma, 2 = subnormal)" | tee -a $out/timing.txt
  { CGQ_GEMV_ARITH=$ar timeout 60 tools/chainbench chain 1 20 | head -1; CGQ_GEMV_ARITH=$ar timeout 60 tools/chainbench step 96 30
    for s in "4096 4608" "4096 4096" "13696 4096" "4096 27392" "4096 65024"; do CGQ_GEMV_ARITH=$ar timeout 60 tools/chainbench single $s 1 10; done; } 2>&1 | tee -a $out/timing.txt
done
