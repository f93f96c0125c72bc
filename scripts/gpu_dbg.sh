#!/bin/bash
MODE=simple_each timeout 300 python scripts/dbg_m8c.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_decode.py -x -q -m gpu 2>&1 | tail -3
timeout 60 tools/chainbench chain 8 10 | head -1
timeout 60 tools/chainbench chain 1 20 | head -1
for s in "4096 65024" "4096 27392" "4096 4608" "4096 4096" "13696 4096"; do timeout 60 tools/chainbench single $s 1 10; done
timeout 60 tools/chainbench step 96 30
