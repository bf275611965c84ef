#!/bin/bash
# correctness (ops + unet + fullsize) then timeline + bench with per-op dump
mkdir -p gpurun_out
R=${1:-x}
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > gpurun_out/pytest_$R.log
echo "== pytest: $(tail -1 gpurun_out/pytest_$R.log)"
python tools/igemm_timeline.py 2>&1 | tail -9
timeout 300 python bench.py --dump-ops gpurun_out/ops_$R.csv --no-cpu-baseline > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
echo "== bench: $(python -c "import json;d=json.load(open('gpurun_out/bench_$R.json'));print(d['value'], d['ms_per_step'], d['roofline']['by_kernel_ms'])" 2>&1 | tail -1)"
tail -2 gpurun_out/bench_$R.err
