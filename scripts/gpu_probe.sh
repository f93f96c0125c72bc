#!/bin/bash
# design probe of the one-launch decode step (tools/mkprobe.cu): box width, barrier cost, consumer headroom, ring depth
tag=${1:-probe}; out=gpurun_out/$tag; mkdir -p $out
make -C tools mkprobe > /dev/null 2>&1
P=tools/mkprobe
{
echo "# pure streaming, no barrier, no compute";
for cfg in "32 1" "64 1" "64 2" "128 1" "128 2" "128 4"; do timeout 60 $P $cfg 20 0 0; done
echo "# + grid barrier between linears";
for cfg in "32 1" "64 2" "128 4"; do timeout 60 $P $cfg 20 0 1; done
echo "# + consumer cost (clocks per stage per team; 4 teams)";
for sp in 400 800 1200 1600; do timeout 60 $P 32 1 20 $sp 1; done
for sp in 800 1600; do timeout 60 $P 64 2 20 $sp 1; timeout 60 $P 128 4 20 $sp 1; done
echo "# ring depth (BW=32, barrier, spin 800)";
for s in 6 10 14; do timeout 60 $P 32 1 $s 800 1; done
} 2>&1 | tee $out/mkprobe.txt
