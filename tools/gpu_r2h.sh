#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -s -x -k "attention" 2>&1 | tail -70 > gpurun_out/r2h_pytest_attn.log; echo "== attn tests: $(tail -1 gpurun_out/r2h_pytest_attn.log)"
grep -E "FAILED|Error|error" gpurun_out/r2h_pytest_attn.log | head -20
grep -E "poly=(4|12|13)" gpurun_out/r2h_pytest_attn.log | head -30
timeout 600 python tools/attn_bench.py gpurun_out/r2h_attn_bench.json 2>&1 | tail -8
