#!/bin/bash
# round 2, call 4: step program timeline + diagnosis switches + unit tests
tag=${1:-mk3}; out=gpurun_out/$tag; mkdir -p $out
make -C tools > /dev/null 2>&1
{
CGQ_STEP_TRACE=1 timeout 120 tools/chainbench mkstep 96 20
echo "== no MMA work"; CGQ_STEP_DBG=1 timeout 120 tools/chainbench mkstep 96 20 | tail -1
echo "== no prologue"; CGQ_STEP_DBG=2 timeout 120 tools/chainbench mkstep 96 20 | tail -1
echo "== neither"; CGQ_STEP_DBG=3 CGQ_STEP_TRACE=1 timeout 120 tools/chainbench mkstep 96 20 | tail -8
timeout 120 tools/chainbench mk 20
} 2>&1 | tee $out/mk.txt
timeout 900 python -m pytest tests/test_gpu_fused_decode.py -x -q -m gpu 2>&1 | tail -5 | tee $out/fused_tests.txt
CGQ_DBG_OPS=12 timeout 200 compute-sanitizer --print-limit 5 --tool racecheck tools/chainbench mk 1 > $out/mk_racecheck.txt 2>&1
echo "mk racecheck: $(grep -E 'RACECHECK SUMMARY' $out/mk_racecheck.txt | tail -1)"
CGQ_DBG_OPS=12 timeout 200 compute-sanitizer --print-limit 5 --tool synccheck tools/chainbench mk 1 > $out/mk_synccheck.txt 2>&1
echo "mk synccheck: $(grep -E 'ERROR SUMMARY' $out/mk_synccheck.txt | tail -1)"
