#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -s -x -k "not large_mean" 2>&1 | tail -70 > gpurun_out/r2c_pytest_ops.log; echo "== ops tests: $(tail -1 gpurun_out/r2c_pytest_ops.log)"
grep -E "FAILED|Error|error|dynamic range" gpurun_out/r2c_pytest_ops.log | head -20
for args in "1024 1024 20 0" "1024 1024 20 4" "4096 4096 10 4" "1024 77 20 4"; do
  timeout 120 python tools/attn_timeline.py $args > gpurun_out/r2c_tl_$(echo $args | tr ' ' '_').txt 2>&1
done
head -12 gpurun_out/r2c_tl_1024_1024_20_0.txt
head -12 gpurun_out/r2c_tl_1024_1024_20_4.txt
sed -n 20,32p gpurun_out/r2c_tl_4096_4096_10_4.txt
head -8 gpurun_out/r2c_tl_1024_77_20_4.txt
timeout 600 python tools/attn_bench.py gpurun_out/r2c_attn_bench.json 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print("== bench ms/step", d["ms_per_step"], d["roofline"]["by_kernel_ms"])
PY
tail -3 gpurun_out/r2c_bench.err
