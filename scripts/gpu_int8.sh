#!/bin/bash
out=gpurun_out/${1:-int8}; mkdir -p $out
timeout 300 python scripts/time_int8.py 2>&1 | tee $out/time.log | tail -20
