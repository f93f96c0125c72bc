#!/bin/bash
out=gpurun_out/exp5; mkdir -p $out
timeout 120 python scripts/dbg_umma.py 4096 256 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "int4_decode_shapes or big_shapes or properties_full_size_int4 or golden" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -5 $out/pytest.log
{
echo "== chain (umma kernel)"; timeout 60 tools/chainbench chain 1 20
for s in "4096 4096" "4096 4608" "4096 27392" "13696 4096" "4096 65024"; do timeout 60 tools/chainbench single $s 1 10; done
echo "== trace"; timeout 60 tools/chainbench trace 1 | head -24
for cfg in "3 4 2" "4 3 2" "5 2 2" "2 4 2" "3 2 3" "6 4 1"; do set -- $cfg; echo "== STAGES=$1 SLOTS=$2 CPS=$3"; CGQ_UMMA_STAGES=$1 CGQ_UMMA_SLOTS=$2 CGQ_UMMA_CTAS_PER_SM=$3 timeout 60 tools/chainbench chain 1 20 | head -1; CGQ_UMMA_STAGES=$1 CGQ_UMMA_SLOTS=$2 CGQ_UMMA_CTAS_PER_SM=$3 timeout 60 tools/chainbench single 4096 65024 1 10; done
} > $out/log.txt 2>&1
cat $out/log.txt
