#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list + full captures of the top kernels.
mkdir -p gpurun_out
R=${1:-r1}
timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 | tail -60 > gpurun_out/pytest_gpu_$R.log
echo "== pytest: $(tail -1 gpurun_out/pytest_gpu_$R.log)"
timeout 900 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
echo "== bench: $(head -c 600 gpurun_out/bench_$R.json)"
tail -5 gpurun_out/bench_$R.err
timeout 300 python tools/vae_bench.py 1 gpurun_out/ops_vae_$R.csv > gpurun_out/vae_bench_$R.json 2>/dev/null
echo "== vae: $(cat gpurun_out/vae_bench_$R.json)"
if [ "$2" != "noncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv \
  --log-file gpurun_out/launches_$R.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$R.log 2>&1
echo "== launches: $(wc -l < gpurun_out/launches_$R.csv) lines"
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:igemm_pair_kernel -s 300 -c 4 \
  -o gpurun_out/prof_igemm_$R -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_igemm_$R.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:attention_kernel -s 20 -c 2 \
  -o gpurun_out/prof_attn_$R -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn_$R.log 2>&1
ls -la gpurun_out/*.ncu-rep
fi
