#!/bin/bash
mkdir -p gpurun_out
for args in "1024 1024 20 0" "1024 1024 20 4"; do
  timeout 120 python tools/attn_timeline.py $args > gpurun_out/r2d_tl_$(echo $args | tr ' ' '_').txt 2>&1
  head -7 gpurun_out/r2d_tl_$(echo $args | tr ' ' '_').txt; tail -10 gpurun_out/r2d_tl_$(echo $args | tr ' ' '_').txt
done
