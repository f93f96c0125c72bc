#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list and one full capture of the top kernel.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag>
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
nproc >> $out/gpu.txt; grep -m1 'model name' /proc/cpuinfo >> $out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
cat $out/bench.json; tail -5 $out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 226 -c 113 --csv \
  --log-file $out/launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-micro --no-cpu > $out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:w4_gemv -s 226 -c 5 \
  -o $out/prof_w4_gemv python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-micro --no-cpu > $out/ncu_full.log 2>&1
ls -la $out
