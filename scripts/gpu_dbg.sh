#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fused_decode.py -x -q -m gpu 2>&1 | tail -3
timeout 60 tools/chainbench step 96 30
timeout 60 tools/chainbench step 1024 30
