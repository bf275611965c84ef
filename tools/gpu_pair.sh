#!/bin/bash
mkdir -p gpurun_out
for k in linear geglu conv2d; do
  timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k $k -x 2>&1 | tail -25 > gpurun_out/pair_ops_$k.log
  echo "== $k: $(tail -1 gpurun_out/pair_ops_$k.log)"
done
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pair_unet.log
echo "== unet: $(tail -1 gpurun_out/pair_unet.log)"
for pr in 1 0; do
  SDXL_B200_PAIR=$pr timeout 300 python bench.py --steps 16 --dump-ops gpurun_out/ops_pair$pr.csv --no-cpu-baseline > gpurun_out/bench_pair$pr.json 2> gpurun_out/bench_pair$pr.err
  echo "== pair=$pr: $(python -c "import json;d=json.load(open('gpurun_out/bench_pair$pr.json'));print(d['value'], d['ms_per_step'], d['roofline']['by_kernel_ms']['igemm_tcgen05'])" 2>&1 | tail -1)"
  tail -2 gpurun_out/bench_pair$pr.err
done
