#!/bin/bash
# compute-sanitizer over the round-2 decode path after the integer-MMA arithmetic / st.async tail / new attention:
# fused token step (chainbench step) and scripts/sanitize_kernels.py, racecheck / synccheck / memcheck
tag=${1:-san2}; out=gpurun_out/$tag; mkdir -p $out
make -C tools chainbench > /dev/null 2>&1
S="compute-sanitizer --print-limit 30 --launch-timeout 0"
run() { # name tool cmd...
  local name=$1 tool=$2; shift 2
  timeout 420 $S --tool $tool "$@" > $out/${name}_${tool}.txt 2>&1
  echo "$name $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${name}_${tool}.txt | tail -1)"
}
for tool in racecheck synccheck memcheck; do
  run step $tool tools/chainbench step 96 1
  run step1000 $tool tools/chainbench step 1000 1
  run kernels $tool python scripts/sanitize_kernels.py
done 2>&1 | tee $out/summary.txt
