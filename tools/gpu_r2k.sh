#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=1 --master-addr 127.0.0.1 --master-port 29578 tests/mp/load_broadcast.py > gpurun_out/mp_lb1.log 2>&1; echo "== load_broadcast world 1 rc=$?"; grep -E "rank|OK|Error|error" gpurun_out/mp_lb1.log | tail -8
timeout 2400 python -m pytest tests -q -m gpu -s 2>&1 | tail -150 > gpurun_out/r2k_pytest_gpu.log; echo "== pytest: $(tail -1 gpurun_out/r2k_pytest_gpu.log)"
grep -E "PARITY|FAILED|Error" gpurun_out/r2k_pytest_gpu.log | head -20
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench.json'))
r=d["roofline"]
print("== bench ms/step", d["ms_per_step"], "frac", r["frac"], "vs burst", r["frac_vs_burst"], "whole", r["whole_step"]["frac"], "exec", r["whole_step"]["executed_flops"], r["by_kernel_ms_in_step"], d["e2e"]["value"], d["parity"]["final_latent_rel"] if d.get("parity") else None)
PY
tail -3 gpurun_out/r2k_bench.err
for wl in image refiner inpaint; do timeout 600 python bench.py --workload $wl --steps 3 --warmup 1 > gpurun_out/r2k_bench_$wl.json 2> gpurun_out/r2k_bench_$wl.err; echo "== $wl: $(head -c 300 gpurun_out/r2k_bench_$wl.json)"; tail -2 gpurun_out/r2k_bench_$wl.err; done
