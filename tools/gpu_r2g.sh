#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 3 -c 1 -o gpurun_out/r2g_attn_v8 -f python tools/attn_stall.py 8 4096 4096 10 6 > gpurun_out/r2g_ncu8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 3 -c 1 -o gpurun_out/r2g_attn_v0 -f python tools/attn_stall.py 0 4096 4096 10 6 > gpurun_out/r2g_ncu0.log 2>&1
ls -la gpurun_out/r2g*
