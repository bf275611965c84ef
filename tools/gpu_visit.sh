#!/bin/bash
# Short GPU visit while iterating on a kernel: op / UNet / full-size parity tests, one bench line with the per-op CSV.  bash tools/gpu_visit.sh <tag> [quick]
mkdir -p gpurun_out
R=${1:-v}
if [ "$2" == "quick" ]; then
  timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -q -x 2>&1 | tail -3
else
  timeout 1800 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_vae_gpu.py tests/test_fullsize_parity_gpu.py tests/test_fullsize_gpu.py -q -s 2>&1 | grep -E "PARITY|passed|failed|Error|error" | head -30
fi
timeout 600 python bench.py --steps 31 --warmup 4 --no-cpu-baseline --dump-ops gpurun_out/ops_$R.csv > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$R.json"))
print("== ms/step", round(d["ms_per_step"], 3), "steps/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d["roofline"]["by_kernel_ms_in_step"], d["clocks"])
PY
tail -3 gpurun_out/bench_$R.err
