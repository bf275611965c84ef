#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -s -x -k "attention" 2>&1 | tail -70 > gpurun_out/r2f_pytest_attn.log; echo "== attn tests: $(tail -1 gpurun_out/r2f_pytest_attn.log)"
grep -E "FAILED|Error|error" gpurun_out/r2f_pytest_attn.log | head -20
for args in "1024 1024 20 12" "4096 4096 10 12" "1024 77 20 12"; do
  timeout 120 python tools/attn_timeline.py $args > gpurun_out/r2f_tl_$(echo $args | tr ' ' '_').txt 2>&1
done
head -12 gpurun_out/r2f_tl_1024_1024_20_12.txt
sed -n 20,30p gpurun_out/r2f_tl_4096_4096_10_12.txt
head -6 gpurun_out/r2f_tl_1024_77_20_12.txt
timeout 600 python tools/attn_bench.py gpurun_out/r2f_attn_bench.json 2>&1 | tail -8
