#!/bin/bash
# round-end style validation: build check, smoke(), the whole GPU suite, both bench arms
out=gpurun_out/${1:-final}; mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2>$out/bench_ref.err; echo "ref rc=$?"; cut -c1-200 $out/bench_ref.json
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cut -c1-400 $out/bench.json; tail -2 $out/bench.err
