#!/bin/bash
# one slot, everything new this session: sampler parity + timing, hand-over bit-identity + timing
out=gpurun_out/${1:-combo}; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_sampling.py -q -m gpu > $out/pytest_sampling.log 2>&1; echo "sampling pytest rc=$?"; tail -12 $out/pytest_sampling.log
timeout 100 python scripts/time_sampler.py 2>&1 | tee $out/time_sampler.log | tail -5
CGQ_TEST_HANDOVER=1 timeout 300 python -m pytest tests/test_gpu_fused_decode.py -q -m gpu -k "handover" > $out/pytest_hand.log 2>&1; echo "handover pytest rc=$?"; tail -8 $out/pytest_hand.log
timeout 200 python scripts/time_fused_step.py 2>&1 | tee $out/time_hand.log | tail -5
