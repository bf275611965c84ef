#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fullsize_parity_gpu.py tests/test_fullsize_gpu.py -q -s 2>&1 | grep -E "PARITY|passed|failed|Error|error" | head -20
timeout 600 python bench.py --steps 31 --warmup 4 --no-cpu-baseline --dump-ops gpurun_out/r2n_ops.csv > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
print("== LN fold: ms/step", d["ms_per_step"], d["roofline"]["by_kernel_ms_in_step"], d["gpu_launches"])
PY
tail -3 gpurun_out/r2n_bench.err
SDXL_B200_LN_FOLD=0 timeout 600 python bench.py --steps 31 --warmup 4 --no-cpu-baseline > gpurun_out/r2n_bench_nofold.json 2> gpurun_out/r2n_bench_nofold.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench_nofold.json'))
print("== separate LN: ms/step", d["ms_per_step"], d["roofline"]["by_kernel_ms_in_step"], d["gpu_launches"])
PY
