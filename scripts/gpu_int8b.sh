#!/bin/bash
# int8 path, round 2: cluster/DSMEM M=1 kernel + fused step
tag=${1:-int8r2}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_fused_decode.py tests/test_gpu_parity.py tests/test_gpu_triton_reference.py -x -q -m gpu -k "int8 or s8" 2>&1 | tail -4 | tee $out/tests.txt
timeout 600 python scripts/time_int8.py 2>&1 | grep -v "M=2048" | tee $out/timing.txt
