#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "linear or conv2d or geglu" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_fullsize_gpu.py -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 31 --warmup 4 --no-cpu-baseline --dump-ops gpurun_out/r2m_ops.csv > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_bench.json'))
print("== TMA epilogue: ms/step", d["ms_per_step"], d["roofline"]["by_kernel_ms_in_step"])
PY
SDXL_B200_EPI_TMA=0 timeout 600 python bench.py --steps 31 --warmup 4 --no-cpu-baseline > gpurun_out/r2m_bench_notma.json 2> gpurun_out/r2m_bench_notma.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_bench_notma.json'))
print("== transposing epilogue: ms/step", d["ms_per_step"], d["roofline"]["by_kernel_ms_in_step"])
PY
tail -3 gpurun_out/r2m_bench.err
