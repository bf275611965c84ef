#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_cluster.log
echo "== pytest (default clusters): $(tail -1 gpurun_out/pytest_cluster.log)"
for c in 2x2 1x1 2x1 1x2; do
  SDXL_B200_CLUSTER=$c timeout 300 python bench.py --steps 16 --dump-ops gpurun_out/ops_cl$c.csv --no-cpu-baseline > gpurun_out/bench_cl$c.json 2> gpurun_out/bench_cl$c.err
  echo "== $c: $(python -c "import json;d=json.load(open('gpurun_out/bench_cl$c.json'));print(d['value'], d['ms_per_step'], d['roofline']['by_kernel_ms']['igemm_tcgen05'])" 2>&1 | tail -1)"
  tail -2 gpurun_out/bench_cl$c.err
done
