#!/bin/bash
# tensor-parallel decode on N GPUs of one box: bench.py under torchrun (parity gates inside), then the N=1 line
N=${1:-2}; tag=${2:-tp$N}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_n$N.json 2> $out/bench_n$N.err
echo "rc=$?"; tail -c 3000 $out/bench_n$N.json; tail -5 $out/bench_n$N.err
