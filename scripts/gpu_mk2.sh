#!/bin/bash
# round 2, call 3: step program v2 (epilogue warp, MMA order, relaxed barrier polls, SiLU pairing)
tag=${1:-mk2}; out=gpurun_out/$tag; mkdir -p $out
make -C tools > /dev/null 2>&1
{
for bw in 32; do
  echo "== step program BW=$bw"
  CGQ_STEP_BW=$bw timeout 120 tools/chainbench mk 20
  CGQ_STEP_BW=$bw CGQ_STEP_TRACE=1 timeout 120 tools/chainbench mkstep 96 20
  CGQ_STEP_BW=$bw CGQ_STEP_NO_PAIR=1 timeout 120 tools/chainbench mkstep 96 20 | tail -1
  CGQ_STEP_BW=$bw timeout 120 tools/chainbench mkstep 1000 20 | tail -1
done
for st in 8 14; do echo "== ring depth $st"; CGQ_STEP_STAGES=$st timeout 120 tools/chainbench mkstep 96 20 | tail -1; done
timeout 60 tools/chainbench step 96 20
timeout 60 tools/chainbench step 1000 20
} 2>&1 | tee $out/mk.txt
timeout 900 python -m pytest tests/test_gpu_fused_decode.py -x -q -m gpu 2>&1 | tail -5 | tee $out/fused_tests.txt
echo "== release-race reproducer (round-1 script): plain release vs load-dependent release" | tee $out/rootcause2.txt
CGQ_HACK_PLAIN_RELEASE=1 MODE=simple_each timeout 200 python scripts/stress_decode_kernel.py 2>&1 | tail -4 | tee -a $out/rootcause2.txt
MODE=simple_each timeout 200 python scripts/stress_decode_kernel.py 2>&1 | tail -2 | tee -a $out/rootcause2.txt
timeout 300 compute-sanitizer --print-limit 5 --tool synccheck python scripts/sanitize_kernels.py > $out/kernels_synccheck.txt 2>&1
echo "kernels synccheck: $(grep -E 'ERROR SUMMARY' $out/kernels_synccheck.txt | tail -1)"
CGQ_DBG_OPS=12 timeout 200 compute-sanitizer --print-limit 5 --tool racecheck tools/chainbench mk 1 > $out/mk_racecheck.txt 2>&1
echo "mk racecheck: $(grep -E 'RACECHECK SUMMARY' $out/mk_racecheck.txt | tail -1)"
