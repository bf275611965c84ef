#!/bin/bash
# First GPU bring-up: each test group in its own process (a trapped kernel kills the CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for k in linear geglu conv2d group_norm layer_norm attention timestep randn; do
  timeout 400 python -m pytest tests/test_ops_gpu.py -q -m gpu -k $k -x 2>&1 | tail -40 > gpurun_out/ops_$k.log
  echo "== $k: $(tail -1 gpurun_out/ops_$k.log)"
done
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu -s 2>&1 | tail -80 > gpurun_out/unet.log
echo "== unet: $(tail -1 gpurun_out/unet.log)"
