#!/bin/bash
# tcgen05 prefill kernel: parity of the split-K path + microbench
tag=${1:-tc2}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc or tcgen05 or prefill or int8" 2>&1 | tail -3 | tee $out/tests.txt
for sp in 0; do
  echo "== CGQ_TC_SPLITS=$sp (0 = automatic)" | tee -a $out/micro.txt
  CGQ_TC_SPLITS=$sp timeout 300 python - <<'PY' 2>&1 | tee -a $out/micro.txt
import sys, json, torch
sys.path.insert(0, ".")
import bench
peaks = bench.load_peaks()
rows = bench.microbench(torch, torch.device("cuda"), peaks, seqs=(16, 64, 128, 256, 2048), ns=(4608, 13696, 27392))
for r in rows:
    print({k: r.get(k) for k in ("M", "N", "us", "TFLOPs", "tensor_frac", "hbm_frac", "triton_us")})
PY
done
