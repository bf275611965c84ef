#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:attention_kernel -s 20 -c 2 \
  -o gpurun_out/prof_attn_v2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn_v2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:igemm_pair_kernel -s 300 -c 3 \
  -o gpurun_out/prof_igemm_pair -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_igemm_pair.log 2>&1
ls -la gpurun_out/*.ncu-rep
