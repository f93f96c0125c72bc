#!/bin/bash
out=gpurun_out/exp6; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "int4_decode_shapes or big_shapes or properties_full_size_int4 or golden" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -4 $out/pytest.log
{
for cfg in "5 3 2" "6 3 2" "4 3 2" "6 2 2" "3 2 3" "4 2 3" "3 4 2"; do set -- $cfg; echo "== STAGES=$1 SLOTS=$2 CPS=$3"; CGQ_UMMA_STAGES=$1 CGQ_UMMA_SLOTS=$2 CGQ_UMMA_CTAS_PER_SM=$3 timeout 60 tools/chainbench chain 1 20 | head -1; CGQ_UMMA_STAGES=$1 CGQ_UMMA_SLOTS=$2 CGQ_UMMA_CTAS_PER_SM=$3 timeout 60 tools/chainbench single 4096 65024 1 10; CGQ_UMMA_STAGES=$1 CGQ_UMMA_SLOTS=$2 CGQ_UMMA_CTAS_PER_SM=$3 timeout 60 tools/chainbench single 4096 27392 1 10; CGQ_UMMA_STAGES=$1 CGQ_UMMA_SLOTS=$2 CGQ_UMMA_CTAS_PER_SM=$3 timeout 60 tools/chainbench single 4096 4096 1 10; done
echo "== trace"; timeout 60 tools/chainbench trace 1 | head -24
} > $out/log.txt 2>&1
cat $out/log.txt
