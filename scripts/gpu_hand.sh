#!/bin/bash
# experimental tile-granular hand-over: bit-identity test + timing against the default protocol
out=gpurun_out/${1:-hand}; mkdir -p $out
CGQ_TEST_HANDOVER=1 timeout 400 python -m pytest tests/test_gpu_fused_decode.py -q -m gpu -k "handover" > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $out/pytest.log
timeout 240 python scripts/time_fused_step.py 2>&1 | tee $out/time.log | tail -6
