#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -s -x 2>&1 | tail -120 > gpurun_out/r2i_pytest_gpu.log; echo "== pytest: $(tail -1 gpurun_out/r2i_pytest_gpu.log)"
grep -E "PARITY|trajectory|FAILED|Error" gpurun_out/r2i_pytest_gpu.log | head -30
timeout 900 python bench.py --steps 31 --warmup 4 --no-cpu-baseline --dump-ops gpurun_out/r2i_ops.csv > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print("== bench ms/step", d["ms_per_step"], d["roofline"]["by_kernel_ms"], d["e2e"]["value"])
PY
tail -3 gpurun_out/r2i_bench.err
