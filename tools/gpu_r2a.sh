#!/bin/bash
# round 2, visit A: pipe microbenchmarks, operator parity (all attention variants), attention timing, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt
timeout 300 ./tools/ubench/pipes > gpurun_out/r2a_pipes.txt 2>&1; echo "== pipes exit $?"; cat gpurun_out/r2a_pipes.txt
timeout 900 python -m pytest tests/test_ops_gpu.py -q -s -x -k "attention" 2>&1 | tail -40 > gpurun_out/r2a_pytest_attn.log; echo "== attn tests: $(tail -1 gpurun_out/r2a_pytest_attn.log)"
grep -E "rel err|FAILED|Error|error" gpurun_out/r2a_pytest_attn.log | head -40
timeout 600 python tools/attn_bench.py gpurun_out/r2a_attn_bench.json 2>&1 | tail -8
timeout 600 python -m pytest tests/test_ops_gpu.py -q -s -k "not attention" 2>&1 | tail -15 > gpurun_out/r2a_pytest_ops.log; echo "== other ops: $(tail -1 gpurun_out/r2a_pytest_ops.log)"; grep -E "group_norm mean|FAILED" gpurun_out/r2a_pytest_ops.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "== bench: $(head -c 400 gpurun_out/r2a_bench.json)"; tail -3 gpurun_out/r2a_bench.err
