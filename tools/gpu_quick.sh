#!/bin/bash
# Quick GPU check: attention + unet parity, bench with per-op dump.
mkdir -p gpurun_out
R=${1:-q}
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_quick_$R.log
echo "== pytest: $(tail -1 gpurun_out/pytest_quick_$R.log)"
timeout 600 python bench.py --dump-ops gpurun_out/ops_$R.csv ${2:---no-cpu-baseline} > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
echo "== bench: $(head -c 300 gpurun_out/bench_$R.json)"
tail -3 gpurun_out/bench_$R.err
