#!/bin/bash
# compute-sanitizer memcheck over the operator-level parity tests and one tiny UNet / sampler run (slow: ~20x).  bash tools/gpu_sanitize.sh <tag>
mkdir -p gpurun_out
R=${1:-san}
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_ops_gpu.py -q -x \
  -k "linear or conv2d or geglu or qkv_attention or group_norm or layer_norm or timestep or randn" > gpurun_out/sanitizer_ops_$R.log 2>&1
echo "ops rc=$?"; tail -5 gpurun_out/sanitizer_ops_$R.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_unet_gpu.py -q -x \
  -k "golden or inpaint" > gpurun_out/sanitizer_unet_$R.log 2>&1
echo "unet rc=$?"; tail -5 gpurun_out/sanitizer_unet_$R.log
for tool in synccheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_ops_gpu.py -q -x \
    -k "linear or conv2d or geglu or qkv_attention or group_norm or layer_norm" > gpurun_out/sanitizer_${tool}_$R.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/sanitizer_${tool}_$R.log
done
