#!/bin/bash
# quick decode-kernel check: parity of everything that touches the M<=8 int4 kernel + timings
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_decode.py -x -q -m gpu -k "int4 or stress or fused or gemv or program or tp_shard or decode" 2>&1 | tail -3
timeout 60 tools/chainbench chain 1 20 | head -1; timeout 60 tools/chainbench step 96 30
for s in "4096 4608" "4096 4096" "13696 4096" "4096 27392" "4096 65024"; do timeout 60 tools/chainbench single $s 1 10; done
