#!/bin/bash
out=gpurun_out/exp7; mkdir -p $out
export CGQ_GEMV_UMMA=1
timeout 120 python scripts/dbg_umma.py 4096 256 2>&1 | tail -12
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "umma or int4_decode_shapes or properties_full_size_int4" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -6 $out/pytest.log
{
for cfg in "8 4 2" "6 4 2" "10 4 2" "8 2 2" "4 4 3" "5 2 4" "16 8 1"; do set -- $cfg; echo "== STAGES=$1 SLOTS=$2 CPS=$3"; CGQ_UMMA_STAGES=$1 CGQ_UMMA_SLOTS=$2 CGQ_UMMA_CTAS_PER_SM=$3 timeout 60 tools/chainbench chain 1 20 | head -1; CGQ_UMMA_STAGES=$1 CGQ_UMMA_SLOTS=$2 CGQ_UMMA_CTAS_PER_SM=$3 timeout 60 tools/chainbench single 4096 65024 1 10; CGQ_UMMA_STAGES=$1 CGQ_UMMA_SLOTS=$2 CGQ_UMMA_CTAS_PER_SM=$3 timeout 60 tools/chainbench single 4096 27392 1 10; CGQ_UMMA_STAGES=$1 CGQ_UMMA_SLOTS=$2 CGQ_UMMA_CTAS_PER_SM=$3 timeout 60 tools/chainbench single 4096 4096 1 10; done
echo "== trace"; timeout 60 tools/chainbench trace 1 | head -24
} > $out/log.txt 2>&1
cat $out/log.txt
