#!/bin/bash
# compute-sanitizer over the decode path: racecheck / synccheck / memcheck on
#   * the fused token step (w4_gemv_kernel with RMSNorm / SiLU-gate prologues, clusters of 8, decode_attn_kernel,
#     decode_begin) -- tools/chainbench step
#   * the M = 8 chain (exact dequant, per-lane L2 activation loads) -- tools/chainbench chain 8
#   * the persistent program kernel(s) -- tools/chainbench program / mk
#   * scripts/sanitize_kernels.py: multi-wave grids at M = 1..8, int8 decode kernel, sampler
# and the root-cause experiment for the M >= 5 divergence of the subnormal-operand variant: the SAME stress loop on
# today's kernel (load-dependent ring release) with CGQ_GEMV_TRICK_MGT1=1.
tag=${1:-san}; out=gpurun_out/$tag; mkdir -p $out
make -C tools chainbench > /dev/null 2>&1
S="compute-sanitizer --print-limit 30 --launch-timeout 0"
run() { # name tool cmd...
  local name=$1 tool=$2; shift 2
  timeout 420 $S --tool $tool "$@" > $out/${name}_${tool}.txt 2>&1
  echo "$name $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${name}_${tool}.txt | tail -1)"
}
for tool in racecheck synccheck memcheck; do
  run step $tool tools/chainbench step 96 1
  run chain8 $tool tools/chainbench chain 8 1
  CGQ_DBG_OPS=9 run program $tool tools/chainbench program 1
  run kernels $tool python scripts/sanitize_kernels.py
  CGQ_GEMV_TRICK_MGT1=1 run kernels_trick_mgt1 $tool python scripts/sanitize_kernels.py
done
# root cause experiment: 200-launch stress, exact (default) and subnormal-operand variant at M > 1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k stress 2>&1 | tail -3 | tee $out/stress_default.txt
CGQ_GEMV_TRICK_MGT1=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k stress 2>&1 | tail -3 | tee $out/stress_trick_mgt1.txt
# does the variant pay?  M = 8 chain with both
timeout 120 tools/chainbench chain 8 20 | head -1 | tee $out/chain8_exact.txt
CGQ_GEMV_TRICK_MGT1=1 timeout 120 tools/chainbench chain 8 20 | head -1 | tee $out/chain8_trick.txt
