#!/bin/bash
# round-2b closing run on the frozen build: smoke, the whole GPU suite, both bench arms, ncu launch list of one token
# step + full-set capture of the decode kernels inside the fused step, per-shape / chain / step timings, attention
tag=${1:-r2b}; out=gpurun_out/$tag; mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_gpu.log
bash scripts/gpu_bench1.sh $tag | tail -12
make -C tools chainbench > /dev/null 2>&1
{ timeout 60 tools/chainbench chain 1 30 | head -1; for c in 96 1000 3000; do timeout 60 tools/chainbench step $c 30; done
  for s in "4096 4608" "4096 4096" "13696 4096" "4096 27392" "4096 65024" "4096 13696"; do timeout 60 tools/chainbench single $s 1 10; done
  timeout 60 tools/chainbench chain 8 20 | head -1
  timeout 100 tools/chainbench steptrace 96 | head -48
  timeout 200 python scripts/time_attention.py 2>/dev/null | grep decode_attn; } > $out/timing.txt 2>&1
head -5 $out/timing.txt
