#!/bin/bash
# sampler: parity suite + timing
out=gpurun_out/${1:-sample}; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_sampling.py -q -m gpu > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $out/pytest.log
timeout 120 python scripts/time_sampler.py 2>&1 | tee $out/time.log | tail -6
