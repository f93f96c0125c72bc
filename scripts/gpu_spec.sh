#!/bin/bash
out=gpurun_out/${1:-spec}; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_fused_decode.py -q -m gpu -k "speculative or matches_unmodified or golden" > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $out/pytest.log
timeout 600 python bench.py --no-micro --no-cpu > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; tail -3 $out/bench.err
python - <<'PY'
import json
d=json.loads(open("" + __import__("sys").argv[1]).read().strip().splitlines()[-1])
e=d["e2e"]; print("value", d["value"], "e2e", e["value"], "speculative", e["fused_step_speculative_next_step"]["value"], "ref sampler", e["fused_step_reference_sampler"]["value"], "dev_us", e["fused_step_device_us"])
PY
