#!/bin/bash
# quick check after a decode-kernel change: int4 decode parity subset + chain / step / single timings
tag=${1:-imma3}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_decode.py -x -q -m gpu -k "int4_decode or imma or stress or fused or properties or tp_shard" 2>&1 | tail -4 | tee $out/pytest.txt
{ timeout 60 tools/chainbench chain 1 20 | head -1; timeout 60 tools/chainbench step 96 30
  for s in "4096 4608" "4096 4096" "13696 4096" "4096 27392" "4096 65024"; do timeout 60 tools/chainbench single $s 1 10; done
  timeout 60 tools/chainbench trace 1 | sed -n 29,63p; } 2>&1 | tee $out/timing.txt
