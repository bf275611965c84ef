#!/bin/bash
# One full GPU visit: parity tests, bench lines (default, reference arm, whole-image workloads), latent-decoder timing, ncu launch
# list of one timed step + `--set full` captures of the tensor-core kernels and of the HBM-bound kernels.  bash tools/gpu_round.sh <tag> [noncu]
mkdir -p gpurun_out
R=${1:-r2}
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 2400 python -m pytest tests -q -m gpu -s 2>&1 | tail -150 > gpurun_out/pytest_gpu_$R.log
echo "== pytest: $(tail -1 gpurun_out/pytest_gpu_$R.log)"; grep -E "PARITY|FAILED" gpurun_out/pytest_gpu_$R.log
timeout 900 python bench.py --dump-ops gpurun_out/ops_$R.csv > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
echo "== bench: $(head -c 400 gpurun_out/bench_$R.json)"; tail -3 gpurun_out/bench_$R.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference_$R.json 2> gpurun_out/bench_reference_$R.err
echo "== reference arm: $(head -c 300 gpurun_out/bench_reference_$R.json)"
for wl in image refiner inpaint; do
  timeout 900 python bench.py --workload $wl --steps 8 --warmup 2 > gpurun_out/bench_${wl}_$R.json 2> gpurun_out/bench_${wl}_$R.err
  echo "== $wl: $(head -c 200 gpurun_out/bench_${wl}_$R.json)"
done
timeout 300 python tools/vae_bench.py 1 gpurun_out/ops_vae_$R.csv > gpurun_out/vae_bench_$R.json 2>/dev/null
echo "== vae: $(cat gpurun_out/vae_bench_$R.json)"
timeout 300 python tools/attn_bench.py gpurun_out/attn_bench_$R.json | tail -1
if [ "$2" != "noncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv \
  --log-file gpurun_out/launches_$R.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$R.log 2>&1
echo "== launches: $(wc -l < gpurun_out/launches_$R.csv) lines"
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:igemm_pair_kernel -s 300 -c 4 \
  -o gpurun_out/prof_igemm_$R -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_igemm_$R.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:attention_kernel -s 20 -c 2 \
  -o gpurun_out/prof_attn_$R -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn_$R.log 2>&1
timeout 900 ncu --set full --clock-control none --nvtx --nvtx-include "timed/" -k "regex:gn_stats|gn_apply|layernorm|gemv|conv_in|cfg_ddim|cast_f32|phase_split" -c 24 \
  -o gpurun_out/prof_small_$R -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_small_$R.log 2>&1
ls -la gpurun_out/*_$R.ncu-rep
fi
