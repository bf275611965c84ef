#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -s 2>&1 | tail -150 > gpurun_out/r2j_pytest_gpu.log; echo "== pytest: $(tail -1 gpurun_out/r2j_pytest_gpu.log)"
grep -E "PARITY|trajectory|FAILED|Error|group_norm mean" gpurun_out/r2j_pytest_gpu.log | head -40
timeout 900 python bench.py --steps 31 --warmup 4 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print("== bench ms/step", d["ms_per_step"], d["roofline"]["by_kernel_ms"], d["e2e"]["value"])
PY
tail -3 gpurun_out/r2j_bench.err
