#!/bin/bash
# full GPU suite + bench line (tag = $1)
tag=${1:-full}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1; nproc >> $out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log; tail -8 $out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cat $out/bench.json; tail -3 $out/bench.err
