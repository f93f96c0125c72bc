#!/bin/bash
# N=1 bench lines of both arms + the ncu evidence of the same command (launch list, full capture of the decode kernel)
tag=${1:-b1}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; tail -c 600 $out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"; tail -c 400 $out/bench_ref.json
B="python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-micro --no-cpu --no-int8"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 400 ncu --metrics $M --clock-control none -k regex:w4_gemv -s 565 -c 113 --csv --log-file $out/launches_token.csv $B > $out/ncu_launch.log 2>&1
tail -2 $out/launches_token.csv | cut -c1-200
make -C tools > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:w4_gemv -s 147 -c 4 -o $out/prof_w4_gemv_fused tools/chainbench step 96 1 > $out/ncu_full.log 2>&1
ls -la $out
