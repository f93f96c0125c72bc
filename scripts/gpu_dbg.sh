#!/bin/bash
for m in simple_each; do MODE=$m timeout 300 python scripts/dbg_m8c.py 2>&1 | tail -4; done
TAG=default timeout 300 python scripts/dbg_m8.py 2>&1 | grep -v " 0/12" | tail -8
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_decode.py -x -q -m gpu 2>&1 | tail -4
timeout 60 tools/chainbench chain 8 10 | head -1
timeout 60 tools/chainbench chain 1 20 | head -1
timeout 60 tools/chainbench step 96 30
