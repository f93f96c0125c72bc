#!/bin/bash
# round 2, call 2: root-cause experiment of the M >= 5 divergence, sanitizer re-check after the fixes, design probe,
# first runs of the one-launch step program
tag=${1:-mk1}; out=gpurun_out/$tag; mkdir -p $out
make -C tools > /dev/null 2>&1
echo "== root cause: 200-launch stress, plain (pre-fix) ring release" | tee $out/rootcause.txt
CGQ_HACK_PLAIN_RELEASE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k stress 2>&1 | grep -E "passed|failed|AssertionError|differ" | head -6 | tee -a $out/rootcause.txt
echo "== same, exact dequant variant" | tee -a $out/rootcause.txt
CGQ_HACK_PLAIN_RELEASE=1 CGQ_GEMV_TRICK_MGT1=0 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k stress 2>&1 | grep -E "passed|failed|AssertionError|differ" | head -6 | tee -a $out/rootcause.txt
echo "== default build (load-dependent release, subnormal-operand variant at M > 1)" | tee -a $out/rootcause.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k stress 2>&1 | tail -1 | tee -a $out/rootcause.txt
S="compute-sanitizer --print-limit 10 --launch-timeout 0"
for tool in racecheck synccheck; do
  timeout 400 $S --tool $tool python scripts/sanitize_kernels.py > $out/kernels_$tool.txt 2>&1
  echo "kernels $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/kernels_$tool.txt | tail -1)"
done
for tool in racecheck synccheck memcheck; do
  CGQ_DBG_OPS=5 timeout 150 $S --tool $tool tools/chainbench mk 1 > $out/mk_$tool.txt 2>&1
  echo "mk $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/mk_$tool.txt | tail -1)"
done
bash scripts/gpu_probe.sh $tag
for bw in 32 64 128; do
  echo "== step program BW=$bw" | tee -a $out/mk.txt
  CGQ_STEP_BW=$bw timeout 120 tools/chainbench mk 20 2>&1 | tee -a $out/mk.txt
  CGQ_STEP_BW=$bw CGQ_STEP_TRACE=1 timeout 120 tools/chainbench mkstep 96 20 2>&1 | tee -a $out/mk.txt
done
timeout 60 tools/chainbench chain 1 20 | head -1 | tee -a $out/mk.txt
timeout 60 tools/chainbench step 96 20 | tee -a $out/mk.txt
