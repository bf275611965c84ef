#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the 2-device tests, then bench.py under torchrun on N ranks (default workload, and the refiner workload = BASELINE config 4).
#   bash tools/gpu_multi.sh <tag> <N>
mkdir -p gpurun_out
R=${1:-mg}; N=${2:-2}
timeout 900 python -m pytest tests/test_pipeline_gpu.py -q -k "two_gpus or two_devices or two_contexts" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --no-cpu-baseline \
  > gpurun_out/bench_${N}gpu_$R.json 2> gpurun_out/bench_${N}gpu_$R.err
echo "== step x$N: $(head -c 330 gpurun_out/bench_${N}gpu_$R.json)"; tail -2 gpurun_out/bench_${N}gpu_$R.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --workload refiner --steps 6 --warmup 2 \
  > gpurun_out/bench_refiner_${N}gpu_$R.json 2> gpurun_out/bench_refiner_${N}gpu_$R.err
echo "== refiner x$N: $(head -c 330 gpurun_out/bench_refiner_${N}gpu_$R.json)"; tail -2 gpurun_out/bench_refiner_${N}gpu_$R.err
