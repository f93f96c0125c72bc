#!/bin/bash
out=gpurun_out/${1:-prefill}; mkdir -p $out
timeout 200 python scripts/time_prefill.py 2>&1 | grep -v Warning | tee $out/time.log | tail -5
