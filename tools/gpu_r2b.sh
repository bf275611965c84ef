#!/bin/bash
mkdir -p gpurun_out
for args in "1024 1024 20 0" "4096 4096 10 0" "1024 77 20 0"; do
  timeout 120 python tools/attn_timeline.py $args > gpurun_out/r2b_tl_$(echo $args | tr ' ' '_').txt 2>&1
done
head -30 gpurun_out/r2b_tl_1024_1024_20_0.txt
head -45 gpurun_out/r2b_tl_4096_4096_10_0.txt | tail -25
head -12 gpurun_out/r2b_tl_1024_77_20_0.txt
