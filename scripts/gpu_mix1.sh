#!/bin/bash
tag=${1:-mix1}; out=gpurun_out/$tag; mkdir -p $out
make -C tools chainbench > /dev/null 2>&1
{ for s in 5 6 7 8; do echo "STAGES_SMALL=$s"; CGQ_GEMV_STAGES_SMALL=$s timeout 60 tools/chainbench chain 1 20 | head -1; CGQ_GEMV_STAGES_SMALL=$s timeout 60 tools/chainbench step 96 30; done
  for s in 6 8 10 12; do echo "STAGES_2PERSM=$s"; CGQ_GEMV_STAGES_2PERSM=$s timeout 60 tools/chainbench chain 1 20 | head -1; CGQ_GEMV_STAGES_2PERSM=$s timeout 60 tools/chainbench step 96 30; done
} 2>&1 | tee $out/stages.txt
timeout 900 compute-sanitizer --print-limit 30 --launch-timeout 0 --tool racecheck -c 45 tools/chainbench step 96 1 > $out/step_racecheck_first45.txt 2>&1
tail -3 $out/step_racecheck_first45.txt
